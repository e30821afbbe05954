"""Developer probe: the fused sweep with the sequential sum inside (papr_exact.cu) at bench size: time,
path taken, sum bits against the host path's (exact) sum."""
import os, struct, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dtv_utils_b200 as pb

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
n = 1 << log2n
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
torch.cuda.synchronize()
for graph in (False, True):
    for exact in (0, -1):
        eng.set("exact_sum", exact)
        best, dev = 1e9, 1e9
        for _ in range(6):
            t0 = time.perf_counter()
            r = eng.analyze_device(d, n, graph)
            best = min(best, time.perf_counter() - t0)
            dev = min(dev, r.device_ms)
        print(f"n=2^{log2n} graph={int(graph)} exact_sum={exact:2d} wall_ms={best*1e3:.3f} device_ms={dev:.3f} scan_ms={r.scan_ms:.3f} "
              f"launches={r.kernel_launches} sum_path={r.sum_path & 255} why={r.sum_path >> 8} miss={r.fused_miss} "
              f"sum={struct.pack('>d', r.stats.sum).hex()}", flush=True)
if log2n <= 31:
    pinned = torch.empty(2 * n, dtype=torch.float32, pin_memory=True)
    pinned.copy_(d)
    torch.cuda.synchronize()
    eng.set("exact_sum", -1)
    h = eng.analyze_host(pinned, graph=False)
    print("host path sum=" + struct.pack('>d', h.stats.sum).hex(), "sum_path", h.sum_path, flush=True)
