#!/bin/bash
# 8-GPU check of the contract bench (in-kernel exchange over NVSwitch peer memory, NUMA binding of the ranks)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m 2>/dev/null | head -12 > $O/topo4.txt; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" >> $O/topo4.txt; free -g | head -2 >> $O/topo4.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 4 --e2e-steps 3 > $O/bench_n4.json 2> $O/bench_n4.err; tail -c 700 $O/bench_n4.json; grep -i "bound to\|NUMA\|error\|Traceback" $O/bench_n4.err | head
