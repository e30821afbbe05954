#!/bin/bash
# instruction count + duration of the fused sweep kernels (one ncu pass; run under gpurun)
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
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"papr_" -s 7 -c 7 --csv --log-file gpurun_out/r02b_quick_${1:-1dB}.csv python /tmp/xt_one.py ${1:+g} > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02b_quick_${1:-1dB}.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index("Kernel Name")][:40], r[h.index("Metric Name")], r[h.index("Metric Value")])
PY
