#!/bin/bash
# One gpurun call of the development loop: parity suite, contract bench, CLI wall-clock on a real
# file, HBM read ceiling.  Everything it writes goes to gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
{
echo "== host"; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" ; free -g | head -2; df -h /dev/shm | tail -1
nvidia-smi --query-gpu=index,name,pci.bus_id,clocks.sm,clocks.max.sm,power.limit --format=csv
nvidia-smi topo -m 2>/dev/null | head -14
} > $O/host.txt 2>&1

echo "== pytest gpu"; 
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu.txt

echo "== bench"
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json

echo "== read ceiling"
timeout 120 tools/_tune/read_peak 2>&1 | tee $O/read_peak.txt

echo "== CLI on a real file"
timeout 900 python tools/cli_timing.py 2>&1 | tee $O/cli_timing.txt
