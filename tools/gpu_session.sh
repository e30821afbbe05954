#!/bin/bash
# what a 1-GPU gpurun call of the development loop runs (tests, smoke, bench); outputs land in gpurun_out/
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
echo "== smoke"
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json
