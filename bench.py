#!/usr/bin/env python
"""bench.py — the PAPR/CCDF hot path on N B200s of one node (contract: see the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA engine via the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU `papr`

A "step" = one complete analysis (statistics + CCDF of every sample) of this rank's shard of a
synthetic capture: 2^31 I/Q samples = 16 GiB per GPU in the 1 dB mode (BASELINE.json configs[1],
"papr 1 dB-bin CCDF on synthetic IQ, 1xB200", scaled from 4 GiB to north_star's >= 10 GiB).  Weak
scaling: every rank holds its own 16 GiB byte range of one N x 16 GiB capture (SURVEY.md App. A
generator, seed 1, rank r = samples [r*2^31, (r+1)*2^31)).

value   whole-job Gsamples/s with the shard resident in HBM when the timed region starts
e2e     same metric through papr_analyze_host with the capture in pinned HOST memory (H2D inside)
One JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "iq_gsamples_per_s"
UNIT = "Gsamples/s"
BYTES_PER_SAMPLE = 8  # algorithmic bytes per unit: each sample read from HBM once


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, stream copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as ex:  # pragma: no cover
            log("clock sampling unavailable:", ex)
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---- the reference arm ------------------------------------------------------------------------------
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "papr")
    return p if os.path.exists(p) else None


def make_host_sample(nsamples, seed, path):
    """Appendix-A capture on the host (oracle C twin) -> file on tmpfs."""
    import numpy as np
    import oracle_binding
    with open(path, "wb") as f:
        step = 1 << 22
        for k in range(0, nsamples, step):
            f.write(oracle_binding.siggen(k, min(step, nsamples - k), seed).tobytes())


def time_cpu(path, nsamples, graph, repeats=1):
    """Best-of wall time of the reference CPU implementation on `path`; returns (seconds, kind, stdout)."""
    ref = reference_binary()
    best, out = None, b""
    for _ in range(repeats):
        if ref:
            t0 = time.perf_counter()
            r = subprocess.run([ref] + (["-g"] if graph else []) + [path], capture_output=True)
            dt = time.perf_counter() - t0
            out = r.stdout
        else:  # the oracle port (same algorithm restated; histogram formulation, so it is FASTER)
            import oracle_binding
            img = open(path, "rb").read()
            t0 = time.perf_counter()
            out = oracle_binding.run_image(img, graph)
            dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, ("reference" if ref else "port"), out


def host_info(value):
    """host CPU model and core count beside the 1-thread figure, plus the hypothetical rate if every
    core ran its own copy of the (single-threaded) tool on its own byte range - labelled as what it is."""
    model = None
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    n = os.cpu_count() or 1
    return {"host_cpus": n, "host_cpu_model": model,
            "hypothetical_all_cores": {"value": value * n, "unit": UNIT,
                                       "note": "1-thread rate x host_cpus; the reference tool has no threads"}}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = args.steps + args.warmup
    budget = 150.0 / max(total, 1)                      # seconds of CPU work per step
    ns = int(min(1 << 28, max(1 << 20, budget * 20e6)))  # reference ~ 20-28 Msamples/s in 1 dB mode
    ns = 1 << (ns.bit_length() - 1)
    path = "/dev/shm/papr_b200_ref_%d.cfile" % os.getpid()
    try:
        make_host_sample(ns, 1, path)
        for _ in range(args.warmup):
            time_cpu(path, ns, args.graph)
        t0 = time.perf_counter()
        kind = "reference"
        for _ in range(args.steps):
            _, kind, _ = time_cpu(path, ns, args.graph)
        dt = (time.perf_counter() - t0) / max(args.steps, 1)
    finally:
        if os.path.exists(path):
            os.unlink(path)
    v = ns / dt / 1e9
    sample = "first 2^%d samples (%d MiB) of the seed-1 capture, file on tmpfs, 1 thread (the tool is single-threaded)" % (
        ns.bit_length() - 1, ns * 8 >> 20)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1 << args.log2_samples),
            "cpu_baseline": dict({"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample}, **host_info(v)),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n):
    which = "configs[2]" if args.graph else "configs[1] scaled to north_star's >=10 GiB"
    if args.signal == "ofdm32k":
        which = "configs[4]"
    return {"workload": "papr %s CCDF on %d GiB synthetic IQ per GPU (BASELINE %s)"
                        % ("-g 0.1 dB-bin" if args.graph else "1 dB-bin", n * 8 >> 30, which),
            "samples_per_gpu": n, "bytes_per_gpu": n * 8, "graph": bool(args.graph),
            "generator": "SURVEY Appendix A, seed 1" if args.signal == "appendixA" else
                         "DVB-T2-like 32K OFDM, 256-QAM, GI 1/128, x0.2 (dtv_utils_b200.producers)", "l2_policy": "inputs_larger_than_l2 (%d GiB vs 126 MB)" % (n * 8 >> 30),
            "sharding": "byte-range, one shard per rank"}


def bind_to_gpu_numa_node(local):
    """Pin this rank (and every thread it creates later: the staging helpers, CUDA's own) to the CPUs
    next to its GPU before any pinned host memory is allocated, so that the e2e leg's H2D source pages
    sit on the GPU's NUMA node.  Matters at N=8 (every rank streams 16 GiB/step through its own PCIe
    link; pages on the wrong socket cross the inter-socket link); best effort, silent when there is a
    single node or NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:  # NVML ignores CUDA_VISIBLE_DEVICES: translate the ordinal when it is a plain index list
            ids = [x.strip() for x in vis.split(",")]
            if local < len(ids) and ids[local].isdigit():
                h = pynvml.nvmlDeviceGetHandleByIndex(int(ids[local]))
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = len(os.sched_getaffinity(0))
        if after != before:
            log("rank on GPU %d bound to %d of %d CPUs (GPU-local NUMA node)" % (local, after, before))
    except Exception as ex:  # pragma: no cover
        log("NUMA binding skipped:", ex)


# ---- our arm ------------------------------------------------------------------------------------------
def scan_traffic(graph, n, chained):
    """DRAM bytes the dominant kernel moves per launch: bytes per sample from the committed `ncu --set full`
    captures of this kernel (profiles/scan_traffic.json: one entry per mode), scaled to this shard."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        e = t["modes"][("tma_" if chained else "ldg_") + ("graph" if graph else "1dB")]
        return e["dram_bytes_per_sample"] * n, e["source"]
    except Exception:
        return None, None


def roofline_of(mres, n, graph, peak, peak_src, world):
    """`roofline` object for the dominant kernel of a measurement (one scan launch per 2^31 samples)."""
    res = mres["res"]
    scan_launch_ms = mres["scan_ms"] / mres["steps"]
    achieved = BYTES_PER_SAMPLE * n / (scan_launch_ms * 1e-3) / 1e9
    fused = res.mode_used == 2
    chained = (res.sum_path & 0xff) == 1
    traffic, tsrc = scan_traffic(graph, n * (1 if fused else 2), chained)
    return {"bound": "hbm",
            "kernel": ("papr_scan_tma_kernel (stats + CCDF + sequential sum, one sweep)" if chained else
                       "papr_scan_kernel<stats,hist>" if fused else "papr_scan_kernel<stats> + papr_scan_kernel<hist>"),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": tsrc, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * n * (1 if fused else 2), "launch_ms": scan_launch_ms,
            "launches_per_step": max(1, -(-n >> 31)) * (1 if fused else 2)}


def check_sharded_parity(pb, torch, dist, eng, d, n, first, graph, rank, world, stdout_p2p):
    """N > 1: (1) at full size the in-kernel (NVLink peer memory) exchange and the NCCL exchange give the same
    text on every rank; (2) a small capture sharded over the ranks gives the oracle's text for the concatenation
    (rank 0 holds the checker; oracle use = checking only) and the same sequential-sum bits."""
    import struct
    out = {}
    p2p = getattr(eng, "_p2p", None)
    eng._p2p = False
    res_nccl = pb.analyze_sharded(eng, d, n, first, graph, mode=pb.papr.MODE_FUSED)
    eng._p2p = p2p
    same = pb.format_result(res_nccl) == stdout_p2p
    ns = 1 << 22
    small = torch.empty(2 * ns, dtype=torch.float32, device="cuda")
    eng.siggen(small, rank * ns, ns, 7)
    ok_small = True
    for g in (False, True):
        r = pb.analyze_sharded(eng, small, ns, rank * ns, g, mode=pb.papr.MODE_FUSED)
        if rank == 0:
            import oracle_binding
            whole = oracle_binding.siggen(0, world * ns, 7)
            ok_small &= pb.format_result(r) == oracle_binding.run_image(whole.tobytes(), g)
            ok_small &= struct.pack("<d", r.stats.sum) == struct.pack("<d", oracle_binding.analyze(whole, False)[0].sum)
    t = torch.tensor([1 if same else 0, 1 if ok_small else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out["p2p_equals_nccl_exchange_full_size"] = bool(t[0].item())
    out["small_capture_equals_oracle"] = bool(t[1].item())
    out["small_capture"] = "%d x 2^22 samples, seed 7, 1 dB and -g, text + sequential-sum bits" % world
    return out


def h2d_ceiling(torch, pinned, d, barrier, reps=2):
    """What plain pinned->device copies of the same buffer reach with every rank copying at once (the ceiling of
    the e2e leg at this N): GB/s of this rank."""
    nb = pinned.numel() * 4
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(reps + 1):
        barrier()
        ev0.record()
        d[:pinned.numel()].copy_(pinned, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        best = ms if best is None else min(best, ms)
    return nb / best / 1e6


def run_ours(args):
    import torch
    import dtv_utils_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        log("note: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = 1 << args.log2_samples
    first = rank * n
    graph = bool(args.graph)
    eng = pb.Engine(local)
    mode = {"auto": 0, "two_pass": 1, "fused": 2}[args.mode]
    n_big = max(n, 1 << 32) if not args.no_configs else n   # one allocation serves the main workload and c3 / c4
    big = torch.empty(2 * n_big, dtype=torch.float32, device="cuda")
    d = big[:2 * n]
    if args.signal == "ofdm32k":  # DVB-T2-like 32K OFDM (BASELINE configs[4]); same base block on every rank
        from dtv_utils_b200.producers import ofdm_capture
        d.copy_(ofdm_capture(n, seed=1 + rank, device="cuda"))
    else:
        eng.siggen(d, first, n, 1)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ext = torch.cuda.ExternalStream(eng.stream)

    def run_resident(buf, ns, first_idx, g, steps, warmup):
        """W untimed + K timed analyses of the resident shard `buf` (times local to this rank)."""
        def one():
            if world == 1:
                return eng.analyze_device(buf, ns, g)
            return pb.analyze_sharded(eng, buf, ns, first_idx, g, mode=pb.papr.MODE_FUSED if mode != 1 else 1)
        for _ in range(warmup):
            r = one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        la = d2 = 0
        sc = dv = 0.0
        t0 = time.perf_counter()
        e0.record(ext)
        for _ in range(steps):
            r = one()
            la += r.kernel_launches
            sc += r.scan_ms
            dv += r.device_ms
            d2 += r.d2h_bytes
        e1.record(ext)
        barrier()
        wall = time.perf_counter() - t0
        ev = e0.elapsed_time(e1)
        # a step ends with a host synchronisation, so wall >= the device span; take the larger (honest) one
        return {"t": max(wall, ev * 1e-3), "ev_ms": ev, "launches": la, "scan_ms": sc, "dev_ms": dv, "d2h": d2,
                "res": r, "steps": steps}

    def reduce_max(m):
        if dist is None:
            return m
        t = torch.tensor([m["t"], m["scan_ms"], m["dev_ms"], m["ev_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m["t"], m["scan_ms"], m["dev_ms"], m["ev_ms"] = t.tolist()
        lt = torch.tensor([m["launches"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        m["launches"] = int(lt.item())
        return m

    eng.set("mode", mode)
    sampler = ClockSampler(local)
    sampler.start()

    # ---- value: shard resident in HBM ---------------------------------------------------------------
    m = reduce_max(run_resident(d, n, first, graph, args.steps, args.warmup))
    res = m["res"]
    stdout_resident = pb.format_result(res)

    # ---- N > 1: is the merged answer right? ------------------------------------------------------------
    sharded_parity = None
    if world > 1:
        sharded_parity = check_sharded_parity(pb, torch, dist, eng, d, n, first, graph, rank, world, stdout_resident)

    # ---- e2e: capture in pinned host memory, H2D inside the timed region ------------------------------
    n_host, e2e_steps, t_e2e, h2d, d2h, pinned, ceil_gbs = n, 0, 1.0, 0, 0, None, None
    if args.e2e_steps > 0:  # (0 only for profiling runs under ncu)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        import psutil
        avail = psutil.virtual_memory().available
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        n_host = n
        while n_host * 8 > 0.45 * avail / local_world and n_host > (1 << 24):
            n_host >>= 1
        pinned = torch.empty(2 * n_host, dtype=torch.float32, pin_memory=True)
        pinned.copy_(d[:2 * n_host])
        torch.cuda.synchronize()
        scratch = big[2 * n:2 * n + 2 * n_host] if n_big >= n + n_host else torch.empty(2 * n_host, dtype=torch.float32, device="cuda")
        ceil_gbs = h2d_ceiling(torch, pinned, scratch, barrier)
        if world == 1:
            def step_host():
                r = eng.analyze_host(pinned, graph=graph)
                return r, r.h2d_bytes, r.d2h_bytes
        else:
            def step_host():  # every rank streams its own shard in; same exchanges as the resident path
                r = pb.analyze_sharded(eng, None, n_host, rank * n_host, graph, host_image=pinned)
                # bytes this rank moved: the shard itself, its statistics / tile sums and runs / level counts
                nt = -(-n_host // 32768)
                return r, n_host * 8, 104 + nt * 24 + 8 * (r.nlevels + 1)
        for _ in range(min(2, args.warmup)):
            hres, _, _ = step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hres, bi, bo = step_host()
            h2d += bi
            d2h += bo
        barrier()
        t_e2e = time.perf_counter() - t0

    # ---- the other BASELINE configurations, briefly (same GPUs, resident) --------------------------------
    subs = {}
    if not args.no_configs and args.signal == "appendixA":
        ks, kw = max(3, min(5, args.steps)), 3

        def sub(name, what, buf, ns, first_idx, g):
            mm = reduce_max(run_resident(buf, ns, first_idx, g, ks, kw))
            r = mm["res"]
            subs[name] = {"workload": what, "graph": g, "samples_per_gpu": ns, "value": world * ns * ks / mm["t"] / 1e9,
                          "unit": UNIT, "ms_per_step": mm["t"] / ks * 1e3, "hbm_pct_of_8tbs": ns * ks / mm["t"] / 1e9 * 8 / 80 / 1,
                          "scan_ms_per_step": mm["scan_ms"] / ks, "fused_miss": int(r.fused_miss), "levels": int(r.nlevels),
                          "exact_sum": (r.sum_path & 0xff) == 1, "sum_path": int(r.sum_path & 0xff),
                          "roofline_frac_of_measured": BYTES_PER_SAMPLE * ns / (mm["scan_ms"] / ks * 1e-3) / 1e9 / measured_peak()[0]}
            return r
        nb = 1 << 32
        if n_big >= nb:
            eng.siggen(big, rank * nb, nb, 2)
            torch.cuda.synchronize()
            sub("c3_graph_32gib", "papr -g on 32 GiB per GPU (BASELINE configs[2]); seed 2", big, nb, rank * nb, True)
            sub("c4_1dB_32gib_per_gpu", "papr 1 dB on 32 GiB per GPU = %d GiB byte-range-sharded over %d GPU(s) (BASELINE configs[3])"
                % (32 * world, world), big, nb, rank * nb, False)
        from dtv_utils_b200.producers import ofdm_capture
        d.copy_(ofdm_capture(n, seed=1 + rank, device="cuda"))
        torch.cuda.synchronize()
        sub("c5_ofdm32k_graph", "papr -g on a DVB-T2-like 32K OFDM capture, %d GiB per GPU (BASELINE configs[4])" % (n * 8 >> 30),
            d, n, first, True)

    sampler.stop_flag = True
    sampler.join(1.0)

    # ---- max over ranks ------------------------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([t_e2e, -(ceil_gbs or 0.0)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e, ceil_min = t[0].item(), -t[1].item()
        ts = torch.tensor([ceil_gbs or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts)
        ceil_sum = ts.item()
    else:
        ceil_min = ceil_sum = ceil_gbs or 0.0

    if rank == 0:
        value = world * n * args.steps / m["t"] / 1e9
        e2e_value = world * n_host * e2e_steps / t_e2e / 1e9 if e2e_steps else None
        peak, peak_src = measured_peak()
        roof = roofline_of(m, n, graph, peak, peak_src, world)
        roof["headline_frac_all_kernels"] = value / world * BYTES_PER_SAMPLE / peak
        chained = (res.sum_path & 0xff) == 1
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": m["t"] / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(args, n), mode=("fused" if res.mode_used == 2 else "two_pass"),
                               parallelism="byte-range x%d" % world,
                               exchange=("none (single shard)" if world == 1 else
                                         "in-kernel over NVLink peer memory" if getattr(eng, "_p2p", False) else "nccl")),
                # papr.c:104: is stats.sum the reference's sequential double sum, bit for bit, on the timed path?
                "exact_sum": bool(chained or (res.sum_path & 0xff) in (2, 3)),
                "sum_path": {0: "fixed-order (not emulated)", 1: "emulated inside the sweep, chained on the device",
                             2: "two-sweep emulation", 3: "emulated while streaming in"}.get(res.sum_path & 0xff),
                "hbm_gbs": value * BYTES_PER_SAMPLE, "hbm_pct_of_8tbs": value * BYTES_PER_SAMPLE / 8000 * 100,
                "device_ms_per_step": m["dev_ms"] / args.steps, "event_ms_per_step": m["ev_ms"] / args.steps,
                "roofline": roof,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(e2e_steps, 1),
                        "d2h_bytes_per_step": d2h // max(e2e_steps, 1), "steps": e2e_steps,
                        "samples_per_gpu": n_host, "host_memory": "pinned",
                        # plain cudaMemcpyAsync of the same pinned buffers, all ranks at once: what the box allows
                        # (a step ends when the slowest rank is done, so the ceiling of the synchronised job is
                        # n_gpus x the slowest rank's rate; the sum of the ranks' own rates is shown beside it)
                        "h2d_ceiling_gbs_slowest_rank": ceil_min, "h2d_ceiling_gbs_sum_of_ranks": ceil_sum,
                        "frac_of_h2d_ceiling": (e2e_value * BYTES_PER_SAMPLE / (world * ceil_min)) if (e2e_value and ceil_min) else None,
                        "frac_of_sum_of_ranks": (e2e_value * BYTES_PER_SAMPLE / ceil_sum) if (e2e_value and ceil_sum) else None},
                "gpu_launches": m["launches"], "clocks": sampler.result(), "fused_miss": int(res.fused_miss)}
        if subs:
            line["configs"] = subs
        if sharded_parity is not None:
            line["sharded_parity"] = all(v for v in sharded_parity.values() if isinstance(v, bool))
            line["sharded_parity_detail"] = sharded_parity
        if world == 1 and pinned is not None and n_host == n:
            # the drop-in path and the timed resident path print the same text and hold the same sum bits
            import struct
            line["resident_stdout_equals_host_path_stdout"] = bool(pb.format_result(hres) == stdout_resident)
            line["resident_sum_bits_equal_host_path_sum_bits"] = bool(struct.pack("<d", hres.stats.sum) == struct.pack("<d", res.stats.sum))
        if world == 1 and not args.no_cpu and pinned is not None:
            line["cpu_baseline"] = cpu_baseline(args, pinned, n_host, graph, eng, pb)
        print(json.dumps(line), flush=True)
    if dist is not None:
        pb.detach_peer_exchange(eng)  # unmap the peers' windows on every rank, barrier, only then tear down
        eng.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline(args, pinned, n_host, graph, eng, pb):
    """The reference's own CPU `papr` (oracle/_ref/papr, unmodified, 1 thread) on a bounded sample of
    the same capture, on this box's host cores — a reported baseline, not the target."""
    ns = min(n_host, 1 << args.cpu_log2_samples)
    path = "/dev/shm/papr_b200_cpu_%d.cfile" % os.getpid()
    try:
        pinned[:2 * ns].numpy().tofile(path)
        dt, kind, out = time_cpu(path, ns, graph)
        same = pb.format_result(eng.analyze_file(path, graph)) == out
    finally:
        if os.path.exists(path):
            os.unlink(path)
    v = ns / dt / 1e9
    return dict({"value": v, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
                 "sample": "first 2^%d samples (%d MiB) of the same capture, file on tmpfs" % (ns.bit_length() - 1, ns * 8 >> 20),
                 "stdout_identical_to_gpu_on_sample": bool(same)}, **host_info(v))


def main():
    # stdout carries exactly one JSON line: keep NCCL's banner / debug output (NCCL_DEBUG=VERSION is set
    # on some boxes) on stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        # this level prints its banner on stdout whatever NCCL_DEBUG_FILE says; INFO/INIT honours the file, so the
        # rank / topology lines a reader of the log looks for still appear - on stderr
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-samples", type=int, default=31, help="samples per GPU (default 2^31 = 16 GiB)")
    ap.add_argument("--graph", action="store_true", help="-g mode (0.1 dB bins)")
    ap.add_argument("--mode", default="auto", choices=["auto", "two_pass", "fused"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-log2-samples", type=int, default=28)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the c3 / c4 / c5 sub-records")
    ap.add_argument("--signal", default="appendixA", choices=["appendixA", "ofdm32k"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("warmup raised to 3 (timing rules)")
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
