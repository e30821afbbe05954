#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== CLI tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli or config1" 2>&1 | tail -3 | tee $O/pytest_cli.txt
bash tools/cli_quick.sh 2>&1 | tee $O/cli_quick.txt
