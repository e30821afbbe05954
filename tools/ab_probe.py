"""Developer probe: device-resident analysis at bench size with a given build of the library (A/B of kernel
changes on the same box): python tools/ab_probe.py [path/to/libpapr_b200.so] [log2n]."""
import os, struct, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dtv_utils_b200 as pb

lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "-" else None
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 31
if lib:
    pb.papr._lib = pb.papr.load_library(os.path.abspath(lib))
n = 1 << log2n
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
torch.cuda.synchronize()
for graph in (False, True):
    dev, scan = [], []
    for _ in range(12):
        r = eng.analyze_device(d, n, graph)
        dev.append(r.device_ms); scan.append(r.scan_ms)
    dev.sort(); scan.sort()
    print(f"{lib or 'tree'} n=2^{log2n} graph={int(graph)} device_ms min {dev[0]:.3f} med {dev[6]:.3f} | scan_ms min {scan[0]:.3f} med {scan[6]:.3f} "
          f"| launches={r.kernel_launches} sum_path={r.sum_path & 255} miss={r.fused_miss} L={r.nlevels} sum={struct.pack('>d', r.stats.sum).hex()} "
          f"c0={r.counts()[0]} cL={r.counts()[-1]}", flush=True)
