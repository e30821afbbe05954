"""Developer timing probe (not the contract bench): device-resident analysis at a few sizes/modes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtv_utils_b200 as pb

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
n = 1 << log2n
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
torch.cuda.synchronize()
for graph in (False, True):
    for mode, name in ((1, "two_pass"), (2, "fused")):
        eng.set("mode", mode)
        for it in range(4):
            t0 = time.perf_counter()
            r = eng.analyze_device(d, n, graph)
            t1 = time.perf_counter()
        gbs = 8 * n / (r.device_ms * 1e-3) / 1e9
        print(f"n=2^{log2n} graph={int(graph)} {name:8s} device_ms={r.device_ms:.3f} scan_ms={r.scan_ms:.3f} "
              f"wall_ms={(t1-t0)*1e3:.3f} launches={r.kernel_launches} miss={r.fused_miss} L={r.nlevels} "
              f"headline={gbs:.0f} GB/s  scan-only={8*n/(r.scan_ms*1e-3)/1e9:.0f} GB/s(per 8B/sample)", flush=True)
