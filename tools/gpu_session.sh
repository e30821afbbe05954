#!/bin/bash
# One gpurun call of the development loop.  Writes to gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
echo "== CLI on a real file"
timeout 600 python tools/cli_timing.py 2>&1 | tee $O/cli_timing.txt
echo "== e2e trace (pinned host memory)"
timeout 600 python tools/e2e_trace.py 2>&1 | grep -v "trace: device buffer\|trace: pinned staging" | tee $O/e2e_trace.txt
echo "== exact sum on a resident shard"
timeout 300 python tools/exact_probe.py 31 2>&1 | tee $O/exact_probe.txt
echo "== ncu: sequential-sum kernels and the hist-only scan (two-pass, exact_sum=1, 8 GiB)"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'papr_seqsum_kernel|papr_tilesum_kernel|papr_scan_kernel' -s 4 -c 4 -f -o $O/r01_hostpath_kernels \
    python tools/exact_probe.py 30 2 > $O/ncu_hostpath.log 2>&1; tail -3 $O/ncu_hostpath.log
echo "== bench (contract), C3, C5"
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.json
timeout 900 python bench.py --graph --log2-samples 32 --steps 10 > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 400 $O/bench_c3.json
timeout 600 python bench.py --graph --signal ofdm32k --steps 10 > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 400 $O/bench_c5.json
