"""ctypes binding of oracle/libpapr_oracle.so — TEST INFRASTRUCTURE (the parity checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class Stats(C.Structure):
    _fields_ = [("n", C.c_int64), ("sum", C.c_double), ("peak", C.c_float), ("peak_idx", C.c_int64),
                ("re_pos", C.c_float), ("im_pos", C.c_float), ("re_neg", C.c_float), ("im_neg", C.c_float),
                ("re_pos_idx", C.c_int64), ("im_pos_idx", C.c_int64),
                ("re_neg_idx", C.c_int64), ("im_neg_idx", C.c_int64)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(ORACLE_DIR, "libpapr_oracle.so")
    src = os.path.join(ORACLE_DIR, "papr_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.papr_oracle_run_buffer.restype = C.c_long
    lib.papr_oracle_run_buffer.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
    lib.papr_oracle_pass1.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Stats)]
    lib.papr_oracle_pass2.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    lib.papr_oracle_levels.restype = C.c_int
    lib.papr_oracle_levels.argtypes = [C.POINTER(Stats), C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_float), C.c_void_p, C.c_int]
    lib.papr_oracle_siggen.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.papr_oracle_stats_init.argtypes = [C.POINTER(Stats)]
    _lib = lib
    return lib


def run_image(img: bytes, graph: bool) -> bytes:
    """Whole-tool stdout of the oracle on a file image."""
    lib = load()
    cap = 1 << 20
    out = C.create_string_buffer(cap)
    buf = (C.c_char * max(len(img), 1)).from_buffer_copy(img if img else b"\0")
    n = lib.papr_oracle_run_buffer(C.cast(buf, C.c_void_p), len(img), int(graph), out, cap)
    assert n >= 0
    return out.raw[:n]


def analyze(iq: np.ndarray, graph: bool):
    """Structured oracle result on a float32 array holding complete I/Q pairs."""
    lib = load()
    iq = np.ascontiguousarray(iq, dtype=np.float32)
    ns = iq.size // 2
    st = Stats()
    lib.papr_oracle_stats_init(C.byref(st))
    lib.papr_oracle_pass1(iq.ctypes.data, ns, C.byref(st))
    avg, papr = C.c_double(), C.c_float()
    L = lib.papr_oracle_levels(C.byref(st), int(graph), C.byref(avg), C.byref(papr), None, 0)
    level = np.zeros(max(L, 1), np.float32)
    counts = np.zeros(max(L, 1), np.int64)
    if L > 0:
        lib.papr_oracle_levels(C.byref(st), int(graph), C.byref(avg), C.byref(papr), level.ctypes.data, L)
        lib.papr_oracle_pass2(iq.ctypes.data, ns, level.ctypes.data, L, counts.ctypes.data)
    return st, avg.value, papr.value, level[:L], counts[:L]


def siggen(first: int, nsamples: int, seed: int) -> np.ndarray:
    out = np.empty(2 * nsamples, np.float32)
    load().papr_oracle_siggen(out.ctypes.data, first, nsamples, seed)
    return out
