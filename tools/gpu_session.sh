#!/bin/bash
# what the last 1-GPU gpurun call of the development loop ran; outputs land in gpurun_out/
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/pytest_gpu.txt
echo "== A/B: HEAD build vs tree"
[ -f tools/_tune/libpapr_head.so ] && timeout 200 python tools/ab_probe.py tools/_tune/libpapr_head.so 2>&1 | tee $O/ab_head.txt
timeout 200 python tools/ab_probe.py - 2>&1 | tee $O/ab_tree.txt
echo "== bench"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json
echo "== ncu instruction counts"
bash tools/ncu_quick.sh 2>&1 | tail -40 | tee $O/ncu_quick_1dB.txt
bash tools/ncu_quick.sh g 2>&1 | tail -40 | tee $O/ncu_quick_g.txt
