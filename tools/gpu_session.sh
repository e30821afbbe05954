#!/bin/bash
# One gpurun call of the development loop.  Writes to gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
echo "== exact sum on a resident shard"
timeout 300 python tools/exact_probe.py 31 2>&1 | tee $O/exact_probe.txt
echo "== CLI on a real file"
timeout 600 python tools/cli_timing.py 2>&1 | tee $O/cli_timing.txt
echo "== ncu: sequential-sum kernels (exact_sum=1, 8 GiB)"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'papr_seqsum_kernel|papr_tilesum_kernel' -s 2 -c 2 -f -o $O/r01_seqsum_kernels \
    python tools/exact_probe.py 30 2 > $O/ncu_seqsum.log 2>&1; tail -3 $O/ncu_seqsum.log
