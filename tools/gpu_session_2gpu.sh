#!/bin/bash
# 2-GPU checks: sharded parity tests, torchrun bench, the CLI sharding one file over two GPUs.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m 2>/dev/null | head -8 > $O/topo2.txt; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" >> $O/topo2.txt
echo "== pytest gpu (sharded)"
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_gpu_2.txt
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 600 $O/bench_n2.json; grep -i "bound to\|NUMA" $O/bench_n2.err | head -4
echo "== CLI, one 16 GiB file over 2 GPUs"
timeout 600 python - <<'PY' 2>&1 | tee $O/cli_2gpu.txt
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import torch
import dtv_utils_b200 as pb
n = 1 << 31
path = "/dev/shm/papr_cli_2gpu.cfile"
eng = pb.Engine(0)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda:0")
eng.siggen(d, 0, n, 1)
with open(path, "wb") as f:
    step = 1 << 27
    for k in range(0, 2 * n, step):
        f.write(d[k:k + step].cpu().numpy().tobytes())
del d; eng.close(); torch.cuda.empty_cache()
try:
    outs = {}
    for ndev in (1, 2, 2):
        env = dict(os.environ, PAPR_B200_STATS="1", PAPR_B200_TRACE="1", PAPR_B200_DEVICES=str(ndev))
        t0 = time.perf_counter()
        r = subprocess.run([pb.cli_path(), path], capture_output=True, env=env)
        dt = time.perf_counter() - t0
        outs[ndev] = r.stdout
        print(f"[devices={ndev}] wall {dt*1e3:.0f} ms rc={r.returncode}\n{r.stderr.decode()}", flush=True)
    print("stdout identical 1 vs 2 GPUs:", outs[1] == outs[2] and len(outs[1]) > 100)
finally:
    os.unlink(path)
PY
