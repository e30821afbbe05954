"""Developer probe: cost of the exact sequential-sum emulation on a device-resident shard
(exact_sum = 1: tile sums + tile runs = two extra sweeps) next to the default tree sum."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dtv_utils_b200 as pb

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 1 << log2n
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
torch.cuda.synchronize()
for exact in (0, 1):
    eng.set("exact_sum", exact)
    for mode, name in ((1, "two_pass"), (2, "fused")):
        eng.set("mode", mode)
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            r = eng.analyze_device(d, n, False)
            best = min(best, time.perf_counter() - t0)
        print(f"n=2^{log2n} exact_sum={exact} {name:8s} wall_ms={best*1e3:.3f} device_ms={r.device_ms:.3f} launches={r.kernel_launches} "
              f"sum={r.stats.sum!r} d2h={r.d2h_bytes}", flush=True)
