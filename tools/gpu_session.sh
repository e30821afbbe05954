#!/bin/bash
# One gpurun call of the development loop: parity suite, host-path probes.  Writes to gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
echo "== CLI on a real file"
timeout 900 python tools/cli_timing.py 2>&1 | tee $O/cli_timing.txt
echo "== e2e trace (pinned host memory)"
timeout 600 python tools/e2e_trace.py 2>&1 | grep -v "^papr_b200 trace: buffers" | tee $O/e2e_trace.txt
