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
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                             "host_cpus": os.cpu_count()},
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
    eng.set("mode", mode)
    if args.signal == "ofdm32k":  # DVB-T2-like 32K OFDM (BASELINE configs[4]); same base block on every rank
        from dtv_utils_b200.producers import ofdm_capture
        d = ofdm_capture(n, seed=1 + rank, device="cuda")
    else:
        d = torch.empty(2 * n, dtype=torch.float32, device="cuda")
        eng.siggen(d, first, n, 1)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if world == 1:
            return eng.analyze_device(d, n, graph)
        return pb.analyze_sharded(eng, d, n, first, graph, mode=pb.papr.MODE_FUSED if mode != 1 else 1)

    ext = torch.cuda.ExternalStream(eng.stream)
    sampler = ClockSampler(local)
    sampler.start()

    # ---- value: shard resident in HBM ---------------------------------------------------------------
    for _ in range(args.warmup):
        res = step_resident()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, scan_ms, dev_ms = 0, 0.0, 0.0
    t0 = time.perf_counter()
    ev0.record(ext)
    for _ in range(args.steps):
        res = step_resident()
        launches += res.kernel_launches
        scan_ms += res.scan_ms
        dev_ms += res.device_ms
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    ev_ms = ev0.elapsed_time(ev1)
    # a step ends with a host synchronisation, so wall >= the device span; take the larger (honest) one
    t_rank = max(wall, ev_ms * 1e-3)
    stdout_resident = pb.format_result(res)

    # ---- e2e: capture in pinned host memory, H2D inside the timed region ------------------------------
    n_host, e2e_steps, t_e2e, h2d, d2h, pinned = n, 0, 1.0, 0, 0, None
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
        if world == 1:
            def step_host():
                return eng.analyze_host(pinned, graph=graph)
        else:
            class _Acc:
                h2d_bytes = 0
                d2h_bytes = 0

            def step_host():  # every rank streams its own shard in; same two exchanges as the resident path
                pb.analyze_sharded(eng, None, n_host, rank * n_host, graph, host_image=pinned)
                _Acc.h2d_bytes, _Acc.d2h_bytes = n_host * 8, 4096
                return _Acc
        for _ in range(min(2, args.warmup)):
            hres = step_host()
        barrier()
        h2d = d2h = 0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hres = step_host()
            h2d += hres.h2d_bytes
            d2h += hres.d2h_bytes
        barrier()
        t_e2e = time.perf_counter() - t0

    sampler.stop_flag = True
    sampler.join(1.0)

    # ---- max over ranks ------------------------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([t_rank, t_e2e, scan_ms, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_rank, t_e2e, scan_ms, dev_ms = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank == 0:
        value = world * n * args.steps / t_rank / 1e9
        e2e_value = world * n_host * e2e_steps / t_e2e / 1e9 if e2e_steps else None
        peak, peak_src = measured_peak()
        scan_launch_ms = scan_ms / args.steps          # dominant kernel: papr_scan_kernel (one launch per step)
        achieved = BYTES_PER_SAMPLE * n / (scan_launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_rank / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(args, n), mode=("fused" if res.mode_used == 2 else "two_pass"),
                               parallelism="byte-range x%d" % world,
                               exchange=("none (single shard)" if world == 1 else
                                         "in-kernel over NVLink peer memory" if getattr(eng, "_p2p", False) else "nccl")),
                "hbm_gbs": value * BYTES_PER_SAMPLE, "hbm_pct_of_8tbs": value * BYTES_PER_SAMPLE / 8000 * 100,
                "device_ms_per_step": dev_ms / args.steps, "event_ms_per_step": ev_ms / args.steps,
                "roofline": {"bound": "hbm", "kernel": "papr_scan_kernel<stats,hist>" if res.mode_used == 2
                             else "papr_scan_kernel<stats> + papr_scan_kernel<hist>",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * n * (1 if res.mode_used == 2 else 2),
                             "launch_ms": scan_launch_ms,
                             "headline_frac_all_kernels": value / world * BYTES_PER_SAMPLE / peak},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(e2e_steps, 1),
                        "d2h_bytes_per_step": d2h // max(e2e_steps, 1), "steps": e2e_steps,
                        "samples_per_gpu": n_host, "host_memory": "pinned"},
                "gpu_launches": launches, "clocks": sampler.result(), "fused_miss": int(res.fused_miss)}
        if world == 1 and pinned is not None and n_host == n:
            # the drop-in path (exact sequential sum) and the timed resident path print the same text
            line["resident_stdout_equals_host_path_stdout"] = bool(pb.format_result(hres) == stdout_resident)
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
    return {"value": ns / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
            "sample": "first 2^%d samples (%d MiB) of the same capture, file on tmpfs" % (ns.bit_length() - 1, ns * 8 >> 20),
            "host_cpus": os.cpu_count(), "stdout_identical_to_gpu_on_sample": bool(same)}


def main():
    # stdout carries exactly one JSON line: keep NCCL's banner / debug output (NCCL_DEBUG=VERSION is set
    # on some boxes) on stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # this level ignores NCCL_DEBUG_FILE
        os.environ["NCCL_DEBUG"] = "WARN"
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
