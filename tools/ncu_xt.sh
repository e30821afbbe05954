#!/bin/bash
# one ncu --set full capture of the fused sweep + chain kernels at bench size (run under gpurun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/xt_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, dtv_utils_b200 as pb
n = 1 << 31
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
torch.cuda.synchronize()
g = len(sys.argv) > 1 and sys.argv[1] == "g"
for _ in range(3):
    r = eng.analyze_device(d, n, g)
print(r.device_ms, r.scan_ms, r.sum_path)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"papr_scan_tma|papr_xt_" -s 4 -c 3 -f -o gpurun_out/r02_xt_${1:-1dB} python /tmp/xt_one.py ${1:+g} 2>&1 | tail -5
