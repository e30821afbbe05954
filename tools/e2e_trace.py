"""Developer probe: where the time goes in papr_analyze_host from pinned memory (the bench's e2e leg)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dtv_utils_b200 as pb

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
n = 1 << log2n
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
pinned = torch.empty(2 * n, dtype=torch.float32, pin_memory=True)
pinned.copy_(d)
torch.cuda.synchronize()
del d
torch.cuda.empty_cache()
for label, knobs in (("default", {}), ("chunk 16 MiB", {"chunk_bytes": 16 << 20}), ("chunk 256 MiB", {"chunk_bytes": 256 << 20}),
                     ("exact_sum off", {"exact_sum": 0}), ("default again", {})):
    eng.set("chunk_bytes", 64 << 20)
    eng.set("exact_sum", -1)
    for k, v in knobs.items():
        eng.set(k, v)
    best = 1e9
    for it in range(4):
        if it == 3:
            os.environ["PAPR_B200_TRACE"] = "1"
        t0 = time.perf_counter()
        r = eng.analyze_host(pinned, graph=False)
        dt = time.perf_counter() - t0
        os.environ.pop("PAPR_B200_TRACE", None)
        if it:
            best = min(best, dt)
    print(f"[{label}] best wall {best*1e3:.2f} ms -> {n/best/1e9:.3f} Gsamples/s ({8*n/best/1e9:.1f} GB/s)  device_ms={r.device_ms:.2f} scan_ms={r.scan_ms:.3f} launches={r.kernel_launches} d2h={r.d2h_bytes}", flush=True)
