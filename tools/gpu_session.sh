#!/bin/bash
# One gpurun call of the development loop.  Writes to gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
echo "== exact sum on a resident shard"
timeout 300 python tools/exact_probe.py 31 2>&1 | tee $O/exact_probe.txt
