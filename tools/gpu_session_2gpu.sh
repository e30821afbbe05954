#!/bin/bash
# 2-GPU checks: sharded parity tests (peer-memory exchange + NCCL path), torchrun bench exit status.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu (sharded)"
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu_2.txt
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --e2e-steps 1 > $O/bench_n2_quick.json 2> $O/bench_n2_quick.err; echo "torchrun rc=$?"; tail -c 300 $O/bench_n2_quick.json; grep -i "error\|Traceback\|warn" $O/bench_n2_quick.err | head -5
