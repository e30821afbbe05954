#!/bin/bash
# what the last 1-GPU gpurun call of the development loop ran; outputs land in gpurun_out/
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1
echo "== bench"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json
echo "== ncu --set full, 1 dB and -g"
bash tools/ncu_xt.sh 2>&1 | tail -3
bash tools/ncu_xt.sh g 2>&1 | tail -3
for m in 1dB g; do
  ncu -i $O/r02_xt_$m.ncu-rep --page raw --csv > $O/r02_xt_${m}_raw.csv 2>/dev/null
done
ls -la $O/*.ncu-rep
echo "== launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bench_n1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --e2e-steps 1 > /dev/null 2>&1
grep -c papr_ $O/launches_bench_n1.csv
echo "== smoke"
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
