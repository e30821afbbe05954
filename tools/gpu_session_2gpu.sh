#!/bin/bash
# 2-GPU checks: sharded parity tests (peer-memory exchange + NCCL path), torchrun bench.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu (sharded)"
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -25 | tee $O/pytest_gpu_2.txt
echo "== bench N=2 (in-kernel exchange)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 500 $O/bench_n2.json; tail -5 $O/bench_n2.err
echo "== bench N=2 (NCCL exchange, for comparison)"
PAPR_B200_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 2 --e2e-steps 0 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; tail -c 300 $O/bench_n2_nccl.json
