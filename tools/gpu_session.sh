#!/bin/bash
# One gpurun call of the development loop.  Writes to gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
echo "== launch list of the bench command (resident leg)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu > $O/bench_under_ncu.log 2>&1; wc -l $O/launches.csv
echo "== CLI on a real file"
timeout 600 python tools/cli_timing.py 2>&1 | tee $O/cli_timing.txt
