"""GPU tests of the byte-range-sharded path over NCCL (stream-ordered stages + in-place collectives on
the engine's buffers).  world_size 1 runs inside this process on any GPU box; world_size 2 launches
torchrun when two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

import fixtures
import oracle_binding

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_ordered_stages_world1(built):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29617", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        eng = built.Engine(0)
        for name, nfl in (("appA_1M", None), ("burst", None), ("zeros_1000", None)):
            f = np.frombuffer(fixtures.image(name), np.float32)
            d = torch.from_numpy(f.copy()).cuda()
            n = f.size // 2
            for graph in (False, True):
                want = oracle_binding.run_image(f.tobytes(), graph)
                for mode in (1, 2):
                    res = built.analyze_sharded(eng, d, n, 0, graph, mode=mode)
                    assert built.format_result(res) == want, (name, graph, mode)
        eng.close()
    finally:
        dist.destroy_process_group()


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import dtv_utils_b200 as pb, oracle_binding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
eng = pb.Engine(int(os.environ["LOCAL_RANK"]))
n = (1 << 22) + 4096            # per rank; ragged vs the 256-sample batch on purpose (+4096 keeps 16 B alignment)
d = torch.empty(2 * n, dtype=torch.float32, device="cuda")
eng.siggen(d, rank * n, n, 5)
whole = oracle_binding.siggen(0, world * n, 5)
ok = True
for graph in (False, True):
    want = oracle_binding.run_image(whole.tobytes(), graph)
    for mode in (1, 2):
        res = pb.analyze_sharded(eng, d, n, rank * n, graph, mode=mode)
        ok &= pb.format_result(res) == want
    pinned = d.cpu().pin_memory()
    res = pb.analyze_sharded(eng, None, n, rank * n, graph, host_image=pinned)
    ok &= pb.format_result(res) == want
print("RANK", rank, "OK" if ok else "MISMATCH", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_two_gpus_match_oracle(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(_WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(w)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2
