#!/bin/bash
# 2-GPU checks: the sharded parity tests (peer-memory exchange + NCCL path), then optionally the torchrun bench
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu (2-GPU tests)"
timeout 480 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -40 | tee $O/pytest_gpu_2.txt
if [ "$1" = "bench" ]; then
  echo "== bench N=2"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 20 --e2e-steps 2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "torchrun rc=$?"; tail -c 300 $O/bench_n2.json; grep -i "error\|Traceback" $O/bench_n2.err | head -5
fi
