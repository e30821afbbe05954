#!/bin/bash
# what the last 1-GPU gpurun call of the development loop ran; outputs land in gpurun_out/
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== bench"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 200 $O/bench_n1.json
echo "== ncu --set full, 1 dB and -g"
bash tools/ncu_xt.sh 2>&1 | tail -2
bash tools/ncu_xt.sh g 2>&1 | tail -2
for m in 1dB g; do
  ncu -i $O/r02_xt_$m.ncu-rep --page raw --csv > $O/r02_xt_${m}_raw.csv 2>/dev/null
  rm -f $O/r02_xt_$m.ncu-rep
done
