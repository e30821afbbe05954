"""Host-side mirror of the reference tool's interface over the C ABI (include/papr_b200.h).

The reference (drmpeg/dtv-utils papr.c) is a C program whose whole interface is
`papr [-g] <infile>` -> stdout; `main()` below is that interface, `Engine` exposes the stages that
the reference's `main` inlines (papr.c:100-129 statistics, :131-141/:164-173 levels, :143-153/:175-185
CCDF counts, :132-135,154-161,186-190 text).  Everything numeric happens in libpapr_b200.so (CUDA);
this file only marshals pointers.  There is no CPU fallback: without the library or a GPU it raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Optional, Sequence

from .build import build, lib_path

MAX_LEVELS = 2048
MODE_AUTO, MODE_TWO_PASS, MODE_FUSED = 0, 1, 2
FLAG_NONFINITE = 1


class PaprError(RuntimeError):
    pass


class PaprStats(C.Structure):
    """papr_stats: state of the reference's first loop (papr.c:36-49) for one contiguous range."""
    _fields_ = [("n", C.c_uint64), ("sum", C.c_double), ("peak", C.c_float), ("re_pos", C.c_float),
                ("im_pos", C.c_float), ("re_neg", C.c_float), ("im_neg", C.c_float), ("flags", C.c_uint32),
                ("peak_idx", C.c_uint64), ("re_pos_idx", C.c_uint64), ("im_pos_idx", C.c_uint64),
                ("re_neg_idx", C.c_uint64), ("im_neg_idx", C.c_uint64)]

    def as_tuple(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


class PaprResult(C.Structure):
    """papr_result: everything the reference prints, plus diagnostics."""
    _fields_ = [("stats", PaprStats), ("avg", C.c_double), ("papr", C.c_float), ("graph", C.c_int32),
                ("nlevels", C.c_int32), ("level", C.c_float * MAX_LEVELS),
                ("level_count", C.c_int64 * MAX_LEVELS), ("mode_used", C.c_int32), ("fused_miss", C.c_int32),
                ("device_ms", C.c_float), ("scan_ms", C.c_float), ("kernel_launches", C.c_uint32),
                ("sum_path", C.c_uint32), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def counts(self):
        return [self.level_count[j] for j in range(self.nlevels)]

    def levels(self):
        return [self.level[j] for j in range(self.nlevels)]


_lib = None

# every symbol include/papr_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = ["papr_abi_version", "papr_engine_create", "papr_engine_destroy", "papr_last_error",
               "papr_engine_stream", "papr_engine_set", "papr_main", "papr_analyze_host",
               "papr_analyze_device", "papr_analyze_file", "papr_stats_device", "papr_stats_host", "papr_stats_merge",
               "papr_levels", "papr_ccdf_device", "papr_fused_presample", "papr_fused_scan",
               "papr_fused_counts", "papr_format", "papr_result_finish", "papr_siggen_device",
               "papr_engine_device_buffer", "papr_shard_presample_async", "papr_shard_scan_async",
               "papr_shard_counts_async", "papr_shard_finish", "papr_multi_create", "papr_multi_destroy",
               "papr_multi_set", "papr_multi_analyze_host", "papr_multi_analyze_file", "papr_multi_last_error",
               "papr_multi_exchange", "papr_seqsum_prepare", "papr_seqsum_runs", "papr_seqsum_chain",
               "papr_xchg_export", "papr_xchg_attach", "papr_xchg_detach", "papr_shard_analyze_p2p",
               "papr_analyze_fd", "papr_multi_analyze_fd", "papr_tr_reduce_device"]
BUF_PRESAMPLE, BUF_LOCAL_STATS, BUF_COUNTS = 0, 1, 2


def load_library(path: Optional[str] = None):
    """dlopen libpapr_b200.so (building it first if the sources are newer) and declare prototypes."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or build()
    if not os.path.exists(p):
        raise PaprError(f"{p} is missing: run __graft_entry__.build() (nvcc, sm_100a)")
    lib = C.CDLL(p)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    lib.papr_abi_version.restype = i32
    lib.papr_engine_create.argtypes = [i32, C.POINTER(vp)]
    lib.papr_engine_destroy.argtypes = [vp]
    lib.papr_engine_destroy.restype = None
    lib.papr_last_error.argtypes = [vp]
    lib.papr_last_error.restype = C.c_char_p
    lib.papr_engine_stream.argtypes = [vp]
    lib.papr_engine_stream.restype = vp
    lib.papr_engine_set.argtypes = [vp, C.c_char_p, C.c_double]
    lib.papr_main.argtypes = [i32, C.POINTER(C.c_char_p)]
    lib.papr_analyze_host.argtypes = [vp, vp, u64, i32, C.POINTER(PaprResult)]
    lib.papr_analyze_device.argtypes = [vp, vp, u64, i32, C.POINTER(PaprResult)]
    lib.papr_analyze_file.argtypes = [vp, C.c_char_p, i32, C.POINTER(PaprResult)]
    lib.papr_analyze_fd.argtypes = [vp, i32, i32, C.POINTER(PaprResult)]
    lib.papr_multi_analyze_fd.argtypes = [vp, i32, i32, C.POINTER(PaprResult)]
    lib.papr_stats_device.argtypes = [vp, vp, u64, u64, C.POINTER(PaprStats)]
    lib.papr_stats_host.argtypes = [vp, vp, u64, u64, C.POINTER(PaprStats), C.POINTER(vp), C.POINTER(u64)]
    lib.papr_stats_merge.argtypes = [C.POINTER(PaprStats), C.POINTER(PaprStats)]
    lib.papr_stats_merge.restype = None
    lib.papr_levels.argtypes = [C.POINTER(PaprStats), i32, C.POINTER(C.c_double), C.POINTER(C.c_float),
                                C.POINTER(C.c_float), i32]
    lib.papr_ccdf_device.argtypes = [vp, vp, u64, C.POINTER(C.c_float), i32, C.POINTER(C.c_int64)]
    lib.papr_fused_presample.argtypes = [vp, vp, u64, i32, C.POINTER(C.c_double)]
    lib.papr_fused_scan.argtypes = [vp, vp, u64, u64, C.POINTER(C.c_double), i32, C.POINTER(PaprStats)]
    lib.papr_fused_counts.argtypes = [vp, C.POINTER(PaprStats), i32, C.POINTER(C.c_int64)]
    lib.papr_format.argtypes = [C.POINTER(PaprResult), C.c_char_p, C.c_size_t]
    lib.papr_format.restype = C.c_long
    lib.papr_result_finish.argtypes = [C.POINTER(PaprResult), i32]
    lib.papr_siggen_device.argtypes = [vp, vp, u64, u64, u64]
    lib.papr_tr_reduce_device.argtypes = [vp, vp, i32, i32, vp, vp, i32, C.c_float, i32, C.c_float, vp, vp]
    lib.papr_engine_device_buffer.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(u64)]
    lib.papr_shard_presample_async.argtypes = [vp, vp, u64, i32]
    lib.papr_shard_scan_async.argtypes = [vp, vp, u64, u64, i32, i32]
    lib.papr_shard_counts_async.argtypes = [vp, vp, i32, i32, i32, vp, u64]
    lib.papr_shard_finish.argtypes = [vp, i32, C.POINTER(PaprResult)]
    lib.papr_multi_create.argtypes = [i32, C.POINTER(i32), C.POINTER(vp)]
    lib.papr_multi_destroy.argtypes = [vp]
    lib.papr_multi_destroy.restype = None
    lib.papr_multi_set.argtypes = [vp, C.c_char_p, C.c_double]
    lib.papr_multi_analyze_host.argtypes = [vp, vp, u64, i32, C.POINTER(PaprResult)]
    lib.papr_multi_analyze_file.argtypes = [vp, C.c_char_p, i32, C.POINTER(PaprResult)]
    lib.papr_multi_last_error.argtypes = [vp]
    lib.papr_multi_last_error.restype = C.c_char_p
    lib.papr_multi_exchange.argtypes = [vp]
    lib.papr_multi_exchange.restype = C.c_char_p
    lib.papr_seqsum_prepare.argtypes = [vp, vp, u64, C.POINTER(C.c_double)]
    lib.papr_seqsum_runs.argtypes = [vp, vp, u64, C.c_double]
    lib.papr_seqsum_chain.argtypes = [vp, vp, u64, C.POINTER(C.c_double)]
    lib.papr_xchg_export.argtypes = [vp, C.c_char_p]
    lib.papr_xchg_attach.argtypes = [vp, i32, i32, C.c_char_p]
    lib.papr_xchg_detach.argtypes = [vp]
    lib.papr_shard_analyze_p2p.argtypes = [vp, vp, u64, u64, i32, C.POINTER(PaprResult)]
    if path is None:
        _lib = lib
    return lib


def _ptr(x) -> int:
    """Device/host address of a torch tensor, numpy array, ctypes buffer or a plain int."""
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    return C.addressof(x)


# ---- host-only stages (no GPU needed) ---------------------------------------------------------------
def merge_stats(parts: Sequence[PaprStats]) -> PaprStats:
    """Fold per-range states in index order (first occurrence wins ties, papr.c:105-126)."""
    lib = load_library()
    acc = PaprStats.from_buffer_copy(bytes(parts[0]))
    for p in parts[1:]:
        lib.papr_stats_merge(C.byref(acc), C.byref(p))
    return acc


def levels(stats: PaprStats, graph: bool):
    """(avg, papr, [levels]) exactly as papr.c:131-141 / :164-173 evaluate them."""
    lib = load_library()
    avg, papr = C.c_double(), C.c_float()
    buf = (C.c_float * MAX_LEVELS)()
    L = lib.papr_levels(C.byref(stats), int(bool(graph)), C.byref(avg), C.byref(papr), buf, MAX_LEVELS)
    return avg.value, papr.value, [buf[j] for j in range(L)]


def format_result(res: PaprResult) -> bytes:
    """The reference's stdout for this result (papr.c:132-135,154-161 / 186-190)."""
    lib = load_library()
    cap = 4096 + 64 * max(res.nlevels, 0)
    out = C.create_string_buffer(cap)
    n = lib.papr_format(C.byref(res), out, cap)
    if n < 0:
        raise PaprError("papr_format failed")
    return out.raw[:n]


def result_from_parts(stats: PaprStats, graph: bool, counts: Sequence[int]) -> PaprResult:
    lib = load_library()
    res = PaprResult()
    res.stats = stats
    lib.papr_result_finish(C.byref(res), int(bool(graph)))
    for j in range(res.nlevels):
        res.level_count[j] = int(counts[j])
    return res


# ---- the engine ------------------------------------------------------------------------------------
class Engine:
    """One GPU's PAPR/CCDF engine (papr_engine).  Not thread-safe; one per GPU per thread."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.papr_engine_create(device, C.byref(h))
        if rc != 0:
            raise PaprError(f"papr_engine_create failed ({rc}): {self.lib.papr_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.papr_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            raise PaprError(f"{what} failed ({rc}): {self.lib.papr_last_error(self.h).decode()}")
        return rc

    def set(self, name: str, value: float):
        self._check(self.lib.papr_engine_set(self.h, name.encode(), float(value)), "papr_engine_set")

    @property
    def stream(self) -> int:
        return self.lib.papr_engine_stream(self.h) or 0

    # whole analysis ------------------------------------------------------------------------------
    def analyze_device(self, d_iq, nsamples: int, graph: bool = False) -> PaprResult:
        res = PaprResult()
        self._check(self.lib.papr_analyze_device(self.h, _ptr(d_iq), nsamples, int(bool(graph)), C.byref(res)),
                    "papr_analyze_device")
        return res

    def analyze_host(self, image, nbytes: Optional[int] = None, graph: bool = False) -> PaprResult:
        """`image` = the capture's file image: bytes, numpy array, torch CPU (pinned) tensor or address."""
        keep = None
        if isinstance(image, (bytes, bytearray)):
            nbytes = len(image) if nbytes is None else nbytes
            keep = (C.c_char * max(len(image), 1)).from_buffer_copy(image if image else b"\0")
            addr = C.addressof(keep)
        else:
            if nbytes is None:
                nbytes = image.numel() * image.element_size() if hasattr(image, "numel") else image.nbytes
            addr = _ptr(image)
        res = PaprResult()
        self._check(self.lib.papr_analyze_host(self.h, addr, nbytes, int(bool(graph)), C.byref(res)),
                    "papr_analyze_host")
        del keep
        return res

    def analyze_file(self, path: str, graph: bool = False) -> PaprResult:
        res = PaprResult()
        self._check(self.lib.papr_analyze_file(self.h, os.fsencode(path), int(bool(graph)), C.byref(res)),
                    "papr_analyze_file")
        return res

    def analyze_fd(self, fd: int, graph: bool = False) -> PaprResult:
        """An open descriptor: a regular file, or a pipe / FIFO (read once, streamed to the GPU as it arrives)."""
        res = PaprResult()
        self._check(self.lib.papr_analyze_fd(self.h, fd, int(bool(graph)), C.byref(res)), "papr_analyze_fd")
        return res

    # stages (sharded callers) --------------------------------------------------------------------
    def stats_shard(self, d_iq, nsamples: int, first_index: int = 0) -> PaprStats:
        st = PaprStats()
        self._check(self.lib.papr_stats_device(self.h, _ptr(d_iq), nsamples, first_index, C.byref(st)),
                    "papr_stats_device")
        return st

    def stats_shard_host(self, image, nbytes: int, first_index: int = 0):
        """-> (stats, device address of the now-resident shard, samples)"""
        st, dp, ns = PaprStats(), C.c_void_p(), C.c_uint64()
        self._check(self.lib.papr_stats_host(self.h, _ptr(image), nbytes, first_index, C.byref(st), C.byref(dp),
                                             C.byref(ns)), "papr_stats_host")
        return st, dp.value or 0, ns.value

    def ccdf_shard(self, d_iq, nsamples: int, level: Sequence[float]):
        L = len(level)
        lv = (C.c_float * max(L, 1))(*level)
        cnt = (C.c_int64 * max(L, 1))()
        self._check(self.lib.papr_ccdf_device(self.h, _ptr(d_iq), nsamples, lv, L, cnt), "papr_ccdf_device")
        return [cnt[j] for j in range(L)]

    def fused_presample(self, d_iq, nsamples: int, graph: bool = False):
        pre = (C.c_double * 4)()
        self._check(self.lib.papr_fused_presample(self.h, _ptr(d_iq), nsamples, int(bool(graph)), pre),
                    "papr_fused_presample")
        return [pre[i] for i in range(4)]

    def fused_scan(self, d_iq, nsamples: int, first_index: int, pre: Sequence[float], graph: bool) -> PaprStats:
        st = PaprStats()
        p = (C.c_double * 4)(*pre)
        self._check(self.lib.papr_fused_scan(self.h, _ptr(d_iq), nsamples, first_index, p, int(bool(graph)),
                                             C.byref(st)), "papr_fused_scan")
        return st

    def fused_counts(self, merged: PaprStats, graph: bool):
        """-> (miss, counts).  miss=True: thresholds fell outside the predicted windows on this shard."""
        cnt = (C.c_int64 * MAX_LEVELS)()
        rc = self._check(self.lib.papr_fused_counts(self.h, C.byref(merged), int(bool(graph)), cnt),
                         "papr_fused_counts")
        return rc == 1, cnt

    # exact sequential sum across shards (papr.c:104; see include/papr_b200.h) ---------------------
    def seqsum_prepare(self, d_iq, nsamples: int) -> float:
        a = C.c_double()
        self._check(self.lib.papr_seqsum_prepare(self.h, _ptr(d_iq), nsamples, C.byref(a)), "papr_seqsum_prepare")
        return a.value

    def seqsum_runs(self, d_iq, nsamples: int, pre: float) -> bool:
        """-> applicable (False: a tile sum is NaN/Inf, so is the reference's sum)"""
        return self._check(self.lib.papr_seqsum_runs(self.h, _ptr(d_iq), nsamples, float(pre)),
                           "papr_seqsum_runs") == 0

    def seqsum_chain(self, d_iq, nsamples: int, state: float) -> float:
        s = C.c_double(state)
        self._check(self.lib.papr_seqsum_chain(self.h, _ptr(d_iq), nsamples, C.byref(s)), "papr_seqsum_chain")
        return s.value

    # exchanges fused into the kernels over peer memory (see include/papr_b200.h) -------------------
    def xchg_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.papr_xchg_export(self.h, buf), "papr_xchg_export")
        return buf.raw

    def xchg_attach(self, rank: int, world: int, handles: bytes):
        self._check(self.lib.papr_xchg_attach(self.h, rank, world, handles), "papr_xchg_attach")

    def xchg_detach(self):
        self._check(self.lib.papr_xchg_detach(self.h), "papr_xchg_detach")

    def shard_analyze_p2p(self, d_iq, nsamples: int, first_index: int, graph: bool) -> PaprResult:
        res = PaprResult()
        self._check(self.lib.papr_shard_analyze_p2p(self.h, _ptr(d_iq), nsamples, first_index, int(bool(graph)),
                                                    C.byref(res)), "papr_shard_analyze_p2p")
        return res

    # stream-ordered stages (see include/papr_b200.h) ---------------------------------------------------
    def device_buffer(self, which: int, dtype):
        """torch tensor aliasing one of the engine's exchange buffers (zero copy)."""
        import torch
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        self._check(self.lib.papr_engine_device_buffer(self.h, which, C.byref(ptr), C.byref(nbytes)),
                    "papr_engine_device_buffer")
        itemsize = torch.empty((), dtype=dtype).element_size()
        typestr = {torch.float64: "<f8", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]

        class _Alias:
            __cuda_array_interface__ = {"shape": (nbytes.value // itemsize,), "typestr": typestr,
                                        "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(_Alias(), device="cuda")

    def shard_presample_async(self, d_iq, nsamples: int, graph: bool):
        self._check(self.lib.papr_shard_presample_async(self.h, _ptr(d_iq), nsamples, int(bool(graph))),
                    "papr_shard_presample_async")

    def shard_scan_async(self, d_iq, nsamples: int, first_index: int, graph: bool, fused: bool):
        self._check(self.lib.papr_shard_scan_async(self.h, _ptr(d_iq), nsamples, first_index, int(bool(graph)),
                                                   int(bool(fused))), "papr_shard_scan_async")

    def shard_counts_async(self, d_all_stats, nparts: int, graph: bool, fused: bool, d_iq, nsamples: int):
        self._check(self.lib.papr_shard_counts_async(self.h, _ptr(d_all_stats), nparts, int(bool(graph)),
                                                     int(bool(fused)), _ptr(d_iq), nsamples),
                    "papr_shard_counts_async")

    def shard_finish(self, graph: bool):
        """-> (miss, result of the whole capture)"""
        res = PaprResult()
        rc = self._check(self.lib.papr_shard_finish(self.h, int(bool(graph)), C.byref(res)), "papr_shard_finish")
        return rc == 1, res

    def tr_reduce(self, symbols, kernel, tones, vclip: float = 3.3, iterations: int = 3, amax: float = 1e30):
        """Tone-reservation PAPR reduction (papr_tr_reduce_device) of `symbols` ([nsym, N] complex64 CUDA tensor,
        corrected in place) with reference kernel `kernel` ([N] complex64) and reserved carriers `tones`.
        Returns (reserved-tone values [nsym, ntones] complex64, iterations [nsym] int32)."""
        import torch
        nsym, n = symbols.shape
        tn = torch.as_tensor(tones, dtype=torch.int32, device=symbols.device).contiguous()
        r = torch.empty((nsym, tn.numel()), dtype=torch.complex64, device=symbols.device)
        it = torch.empty(nsym, dtype=torch.int32, device=symbols.device)
        assert symbols.is_contiguous() and symbols.dtype == torch.complex64 and kernel.dtype == torch.complex64
        torch.cuda.synchronize()
        self._check(self.lib.papr_tr_reduce_device(self.h, symbols.data_ptr(), nsym, n, kernel.data_ptr(), tn.data_ptr(),
                                                   tn.numel(), float(vclip), int(iterations), float(amax), r.data_ptr(),
                                                   it.data_ptr()), "papr_tr_reduce_device")
        return r, it

    def siggen(self, d_out, first_index: int, nsamples: int, seed: int):
        self._check(self.lib.papr_siggen_device(self.h, _ptr(d_out), first_index, nsamples, seed),
                    "papr_siggen_device")


class MultiEngine:
    """One capture sharded by byte range over several GPUs of this box, single process (papr_multi).
    `devices` may repeat a device index (virtual shards on one GPU; counts are then summed on the host)."""

    def __init__(self, devices: Sequence[int]):
        self.lib = load_library()
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = self.lib.papr_multi_create(len(devices), arr, C.byref(h))
        if rc != 0:
            raise PaprError(f"papr_multi_create failed ({rc}): {self.lib.papr_multi_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.papr_multi_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc, what):
        if rc < 0:
            raise PaprError(f"{what} failed ({rc}): {self.lib.papr_multi_last_error(self.h).decode()}")

    def set(self, name: str, value: float):
        self._check(self.lib.papr_multi_set(self.h, name.encode(), float(value)), "papr_multi_set")

    @property
    def exchange(self) -> str:
        return self.lib.papr_multi_exchange(self.h).decode()

    def analyze_host(self, image, nbytes: Optional[int] = None, graph: bool = False) -> PaprResult:
        keep = None
        if isinstance(image, (bytes, bytearray)):
            nbytes = len(image) if nbytes is None else nbytes
            keep = (C.c_char * max(len(image), 1)).from_buffer_copy(image if image else b"\0")
            addr = C.addressof(keep)
        else:
            if nbytes is None:
                nbytes = image.numel() * image.element_size() if hasattr(image, "numel") else image.nbytes
            addr = _ptr(image)
        res = PaprResult()
        self._check(self.lib.papr_multi_analyze_host(self.h, addr, nbytes, int(bool(graph)), C.byref(res)),
                    "papr_multi_analyze_host")
        del keep
        return res

    def analyze_file(self, path: str, graph: bool = False) -> PaprResult:
        res = PaprResult()
        self._check(self.lib.papr_multi_analyze_file(self.h, os.fsencode(path), int(bool(graph)), C.byref(res)),
                    "papr_multi_analyze_file")
        return res


# ---- byte-range sharding over ranks (one process per GPU, torch.distributed) -----------------------
def analyze_sharded(engine, d_iq, nsamples: int, first_index: int, graph: bool, mode: int = MODE_TWO_PASS,
                    group=None, host_image=None, exact_sum: Optional[bool] = None) -> PaprResult:
    """Each rank holds samples [first_index, first_index+nsamples) of one capture; ranks are in index
    order.  Two tiny exchanges, no data-path collective: all-gather of the pass-1 states (merged in
    rank order on every rank, so first occurrences and the level table are identical everywhere) and
    one all-reduce (sum) of the integer level counts.  Returns the whole-capture result on every rank.

    exact_sum: emulate the reference's sequential double sum (papr.c:104) bit for bit across the
    ranks.  Default (None) = on.  Device-resident shards in fused mode get it inside the single sweep
    (papr_shard_analyze_p2p: per-tile runs from the TMA-fed scan, the chains of all ranks exchanged over
    NVLink and walked on every GPU); otherwise tile runs are computed per rank and chained rank after
    rank (world_size tiny broadcasts).  False: any fixed summation order (faster on the NCCL path).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    backend = dist.get_backend(group) if dist.is_initialized() else "none"
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if exact_sum is None:
        exact_sum = True
    partial = None
    if backend == "nccl" and isinstance(engine, Engine) and host_image is None:
        res, done = _analyze_sharded_stream_ordered(engine, d_iq, nsamples, first_index, bool(graph), mode, group, exact_sum)
        if done:
            return res
        partial = res  # everything but the sequential sum (the device chain declined): patched in below
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    graph = bool(graph)

    def allreduce_f64(vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, group=group)
        return t.tolist()

    def gather_stats(st: PaprStats):
        if world == 1:
            return [st]
        raw = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).to(dev)
        out = [torch.empty_like(raw) for _ in range(world)]
        dist.all_gather(out, raw, group=group)
        return [PaprStats.from_buffer_copy(bytes(o.cpu().numpy().tobytes())) for o in out]

    def allreduce_counts(cnt, extra=0):
        vals = [int(c) for c in cnt] + [int(extra)]
        if world == 1:
            return vals[:-1], vals[-1]
        t = torch.tensor(vals, dtype=torch.int64, device=dev)
        dist.all_reduce(t, group=group)
        v = t.tolist()
        return v[:-1], v[-1]

    if partial is not None:
        # statistics, and counts for the levels of the approximate sum, are in hand: chain the sequential sum rank
        # after rank, re-derive the levels, and recount only if one of them moved
        merged = PaprStats.from_buffer_copy(bytes(partial.stats))
        if merged.sum == merged.sum and abs(merged.sum) != float("inf"):
            merged.sum = _chain_sequential_sum(engine, d_iq, nsamples, rank, world, dev, group, merged.sum)
        _avg, _papr, lv = levels(merged, graph)
        if lv == partial.levels():
            counts = partial.counts()
        else:
            counts, _ = allreduce_counts(engine.ccdf_shard(d_iq, nsamples, lv))
        out = result_from_parts(merged, graph, counts)
        out.mode_used, out.fused_miss, out.sum_path = partial.mode_used, partial.fused_miss, partial.sum_path
        return out
    if host_image is not None:  # shard in (pinned) host memory: H2D + pass 1 overlapped, then resident
        mode = MODE_TWO_PASS
        local, d_iq, nsamples = engine.stats_shard_host(host_image, nsamples * 8, first_index)
    elif mode == MODE_FUSED:
        pre = allreduce_f64(engine.fused_presample(d_iq, nsamples, graph))
        local = engine.fused_scan(d_iq, nsamples, first_index, pre, graph)
    else:
        local = engine.stats_shard(d_iq, nsamples, first_index)
    merged = merge_stats(gather_stats(local))
    if exact_sum and merged.sum == merged.sum and abs(merged.sum) != float("inf"):
        merged.sum = _chain_sequential_sum(engine, d_iq, nsamples, rank, world, dev, group, merged.sum)
    _avg, _papr, lv = levels(merged, graph)
    L = len(lv)
    counts = [0] * L
    if L:
        miss = 1
        if mode == MODE_FUSED:
            m, cnt = engine.fused_counts(merged, graph)
            counts, miss = allreduce_counts([cnt[j] for j in range(L)], int(m))
        if miss:  # two-pass mode, or some rank's thresholds fell outside its predicted windows
            counts, _ = allreduce_counts(engine.ccdf_shard(d_iq, nsamples, lv))
    return result_from_parts(merged, graph, counts)


def _chain_sequential_sum(engine, d_iq, nsamples, rank, world, dev, group, fallback):
    """papr.c:104 across ranks: every rank reduces its tiles to parity-transducer runs in parallel
    (papr_seqsum_prepare / _runs); the running sum is then handed from rank to rank in index order
    (papr_seqsum_chain), one 8-byte broadcast per rank.  Returns `fallback` when some rank reports
    that the emulation does not apply."""
    import torch
    import torch.distributed as dist

    approx = engine.seqsum_prepare(d_iq, nsamples)
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = approx
        dist.all_reduce(t, group=group)     # one non-zero contribution per slot: an all-gather
        parts = t.tolist()
    else:
        parts = [approx]
    pre = 0.0
    for q in range(rank):
        pre += parts[q]
    ok = engine.seqsum_runs(d_iq, nsamples, pre)
    if world > 1:
        f = torch.tensor([0 if ok else 1], dtype=torch.int64, device=dev)
        dist.all_reduce(f, group=group)
        ok = int(f.item()) == 0
    if not ok:
        return fallback
    state = 0.0
    for r in range(world):                  # the only sequential part: ~1 ms of host work per rank
        if rank == r:
            state = engine.seqsum_chain(d_iq, nsamples, state)
        if world > 1:
            t = torch.tensor([state], dtype=torch.float64, device=dev)
            dist.broadcast(t, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
            state = float(t.item())
    return state


def attach_peer_exchange(engine: "Engine", group=None) -> bool:
    """Map every rank's exchange window into every other rank (cudaIpc over NVLink) so that
    papr_shard_analyze_p2p can run.  Collective: every rank of `group` must call it.  Returns False
    (on every rank) when any rank could not attach, e.g. no peer access; the NCCL path remains."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ok = world <= 16 and os.environ.get("PAPR_B200_EXCHANGE", "").lower() != "nccl"
    handle = b""
    if ok:
        try:
            handle = engine.xchg_export()
        except PaprError:
            ok = False
    gathered = [None] * world
    dist.all_gather_object(gathered, handle if ok else b"", group=group)
    ok = ok and all(len(h) == 64 for h in gathered)
    if ok:
        try:
            engine.xchg_attach(rank, world, b"".join(gathered))
        except PaprError:
            ok = False
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    agreed = bool(flag.item())
    if agreed and dev == "cuda":
        # prove the windows before relying on them: one tiny collective analysis through the in-kernel
        # exchange (a peer whose stores do not arrive costs a 4 s device-side timeout here, once,
        # instead of an error in the middle of the caller's work)
        probe = torch.zeros(2 * 4096, dtype=torch.float32, device="cuda")
        try:
            engine.shard_analyze_p2p(probe, 4096, rank * 4096, False)
        except PaprError:
            flag.zero_()
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        agreed = bool(flag.item())
    if ok and not agreed:  # some rank could not attach or the probe failed: nobody uses the windows
        engine.xchg_detach()
    return agreed


def detach_peer_exchange(engine: "Engine", group=None) -> None:
    """Collective counterpart of attach_peer_exchange: every rank unmaps its peers' windows, then all
    ranks meet at a barrier, so that no rank frees its window while another still has it mapped.  Call
    it before the engines are closed / the process group is destroyed."""
    import torch.distributed as dist

    if getattr(engine, "_p2p", None):
        engine.xchg_detach()
    engine._p2p = None
    if dist.is_initialized():
        dist.barrier(group=group)


def _analyze_sharded_stream_ordered(engine: "Engine", d_iq, nsamples, first_index, graph, mode, group, exact_sum=True):
    """Device-resident shards, one process per GPU.  Fused mode: ONE library call per rank, the
    exchanges happen inside the kernels over peer memory (papr_shard_analyze_p2p), the sequential sum
    included.  Otherwise (two-pass mode, or peer mapping unavailable) the NCCL path: kernels and three
    tiny collectives enqueued back to back on the engine's stream, in place on the engine's device
    buffers; one host synchronisation - its sum is a fixed-order tree sum, so it only serves callers
    that said exact_sum=False.  Returns (result, done): done=False means the caller still has to supply
    the sequential sum (result = everything else, or None: nothing computed yet); every rank reaches the
    same verdict (the chain's status and a failed exchange are agreed on the device)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    fused = mode != MODE_TWO_PASS
    if fused:
        p2p = getattr(engine, "_p2p", None)
        if p2p is None:
            p2p = engine._p2p = attach_peer_exchange(engine, group)
        if p2p:
            # (no agreement needed: a rank whose shard is too small for the chained sweep declines inside the chain
            # exchange, and then every rank reports the fall-back - sum_path 2 - together)
            try:
                res = engine.shard_analyze_p2p(d_iq, nsamples, first_index, graph)
            except PaprError:
                # a rank did not show up in time: every rank gets the same error (abort words in the windows),
                # so all of them drop the windows together and continue over NCCL
                engine._p2p = False
                res = None
            if res is not None:
                declined = (res.sum_path & 0xff) == 2
                small = (res.sum_path & 0xff) == 0 and res.stats.sum == res.stats.sum and abs(res.stats.sum) != float("inf")
                return res, (not exact_sum or not (declined or small))
    if exact_sum:
        return None, False
    st = getattr(engine, "_xchg", None)
    if st is None:
        st = engine._xchg = {
            "stream": torch.cuda.ExternalStream(engine.stream),
            "pre": engine.device_buffer(BUF_PRESAMPLE, torch.float64),
            "local": engine.device_buffer(BUF_LOCAL_STATS, torch.uint8),
            "counts": engine.device_buffer(BUF_COUNTS, torch.int64),
        }
        st["all"] = torch.empty(world * st["local"].numel(), dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(st["stream"]):
        if fused:
            engine.shard_presample_async(d_iq, nsamples, graph)
            dist.all_reduce(st["pre"], group=group)
        engine.shard_scan_async(d_iq, nsamples, first_index, graph, fused)
        dist.all_gather_into_tensor(st["all"], st["local"], group=group)
        engine.shard_counts_async(st["all"], world, graph, fused, d_iq, nsamples)
        dist.all_reduce(st["counts"], group=group)
        miss, res = engine.shard_finish(graph)
        if miss:  # some rank's thresholds fell outside its predicted windows: exact pass everywhere
            engine.shard_counts_async(st["all"], world, graph, False, d_iq, nsamples)
            dist.all_reduce(st["counts"], group=group)
            miss2, res = engine.shard_finish(graph)
            if miss2:
                raise PaprError("exact CCDF pass reported a miss")
            res.fused_miss = 1
    return res, True


# ---- the process boundary --------------------------------------------------------------------------
def main(argv: Optional[Sequence[str]] = None) -> int:
    """`papr [-g] <infile>` — same argv surface, stdout/stderr text and exit status as papr.c:32-196."""
    lib = load_library()
    args = list(sys.argv if argv is None else argv)
    arr = (C.c_char_p * (len(args) + 1))(*[os.fsencode(a) for a in args], None)
    sys.stdout.flush()
    rc = lib.papr_main(len(args), arr)
    C.CDLL(None).fflush(None)
    return rc
