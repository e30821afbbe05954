#!/bin/bash
# One gpurun call of the development loop: parity suite, contract bench, host-path probes.
# Everything it writes goes to gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu.txt
echo "== e2e trace (pinned host memory)"
timeout 600 python tools/e2e_trace.py 2>&1 | tee $O/e2e_trace.txt
echo "== CLI on a real file"
timeout 900 python tools/cli_timing.py 2>&1 | tee $O/cli_timing.txt
echo "== bench"
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 700 $O/bench_n1.json
