#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_path or seeded or full_size or epilogue or two_launch or bsearch or forced_miss or chunk_boundary or host_buffer" 2>&1 | tail -5 | tee $O/pytest_subset.txt
timeout 90 python tools/ab_probe.py - 2>&1 | tail -2 | tee $O/ab_tree2.txt
timeout 90 python tools/ab_probe.py - 2>&1 | tail -2 | tee -a $O/ab_tree2.txt
