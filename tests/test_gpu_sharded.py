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
        # fused mode went through the peer-memory exchange kernels (world 1: the own window only) ...
        assert eng._p2p is True
        # ... and the NCCL path gives the same text
        eng._p2p = False
        f = np.frombuffer(fixtures.image("appA_1M"), np.float32)
        d = torch.from_numpy(f.copy()).cuda()
        for graph in (False, True):
            res = built.analyze_sharded(eng, d, f.size // 2, 0, graph, mode=2)
            assert built.format_result(res) == oracle_binding.run_image(f.tobytes(), graph)
        eng._p2p = True
        # forced miss (predicted mean off by 5 %): the exact thresholds fall outside the windows -> status
        # word != 0 after the in-kernel sum -> exact pass, still the right text
        big = fixtures.siggen(0, 1 << 22, 9)
        d = torch.from_numpy(big).cuda()
        eng.set("predict_bias", 1.05)
        res = built.analyze_sharded(eng, d, 1 << 22, 0, True, mode=2)
        eng.set("predict_bias", 1.0)
        assert res.fused_miss == 1
        assert built.format_result(res) == oracle_binding.run_image(big.tobytes(), True)
        built.detach_peer_exchange(eng)
        assert eng._p2p is None
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
import struct
whole = oracle_binding.siggen(0, world * n, 5)
seq_sum = oracle_binding.analyze(whole, False)[0].sum
ok = True
def mark(what):
    print("rank", rank, "at", what, "ok" if ok else "MISMATCH so far", flush=True)
for graph in (False, True):
    want = oracle_binding.run_image(whole.tobytes(), graph)
    mark("graph=%%d" %% graph)
    for mode in (1, 2):
        res = pb.analyze_sharded(eng, d, n, rank * n, graph, mode=mode)
        ok &= pb.format_result(res) == want
        # the DEFAULT is the reference's sequential sum over the whole capture, bit for bit, on every rank
        ok &= struct.pack("<d", res.stats.sum) == struct.pack("<d", seq_sum)
    ok &= eng._p2p is True          # mode 2 ran with the exchanges inside the kernels over NVLink peer memory
    # ... and its sum came from the chain built inside the sweep, exchanged over NVLink and walked on every GPU
    res = pb.analyze_sharded(eng, d, n, rank * n, graph, mode=2)
    ok &= (res.sum_path & 0xff) == 1
    if not ok: print("rank", rank, "graph", graph, "sum_path", res.sum_path, flush=True)
    mark("p2p default done")
    eng._p2p = False                # the NCCL path: same text (fixed-order sum when the caller waives exactness)
    ok &= pb.format_result(pb.analyze_sharded(eng, d, n, rank * n, graph, mode=2, exact_sum=False)) == want
    ok &= struct.pack("<d", pb.analyze_sharded(eng, d, n, rank * n, graph, mode=2).stats.sum) == struct.pack("<d", seq_sum)
    eng._p2p = True
    mark("nccl path done")
    eng.set("predict_bias", 1.05)   # forced miss: every rank must take the exact-pass branch together
    res = pb.analyze_sharded(eng, d, n, rank * n, graph, mode=2)
    eng.set("predict_bias", 1.0)
    ok &= res.fused_miss == 1 and pb.format_result(res) == want
    mark("forced miss done")
    # the counts are taken against levels derived from the fixed-order sums while the chain is still running; a
    # (forced) disagreement with the levels of the chained sum must make every rank count again, together
    for bias in (1.0 + 3e-7, 1.25):
        eng.set("epilogue_bias", bias)
        res = pb.analyze_sharded(eng, d, n, rank * n, graph, mode=2)
        eng.set("epilogue_bias", 1.0)
        ok &= res.fused_miss == 2 and pb.format_result(res) == want
        ok &= struct.pack("<d", res.stats.sum) == struct.pack("<d", seq_sum)
        if not ok: print("rank", rank, "graph", graph, "bias", bias, "fused_miss", res.fused_miss, flush=True)
    mark("epilogue bias done")
    pinned = d.cpu().pin_memory()
    res = pb.analyze_sharded(eng, None, n, rank * n, graph, host_image=pinned)
    ok &= pb.format_result(res) == want
    # host-resident shards chain the reference's sequential sum across the ranks: bit for bit
    ok &= struct.pack("<d", res.stats.sum) == struct.pack("<d", seq_sum)
    res = pb.analyze_sharded(eng, d, n, rank * n, graph, mode=2, exact_sum=True)  # and on request for resident ones
    ok &= pb.format_result(res) == want and struct.pack("<d", res.stats.sum) == struct.pack("<d", seq_sum)
# unequal shards, the last one too small for the chained sweep: that rank declines inside the chain exchange, every rank
# reports the fall-back, and the sequential sum is supplied stage by stage - same text, same sum bits
mark("equal shards done")
n1 = 5000
nn = n if rank < world - 1 else n1
tot = (world - 1) * n + n1
d2 = torch.empty(2 * nn, dtype=torch.float32, device="cuda")
eng.siggen(d2, rank * n, nn, 6)
whole2 = oracle_binding.siggen(0, tot, 6)
res = pb.analyze_sharded(eng, d2, nn, rank * n, False, mode=2)
ok &= pb.format_result(res) == oracle_binding.run_image(whole2.tobytes(), False)
ok &= struct.pack("<d", res.stats.sum) == struct.pack("<d", oracle_binding.analyze(whole2, False)[0].sum)
ok &= (res.sum_path & 0xff) == 3 or (res.sum_path & 0xff) == 2   # not the device chain
mark("unequal shards done")
print("RANK", rank, "OK" if ok else "MISMATCH", flush=True)
pb.detach_peer_exchange(eng); eng.close()
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_two_gpus_match_oracle(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(_WORKER % {"root": ROOT})
    try:
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29631", str(w)],
                           capture_output=True, text=True, timeout=420)
    except subprocess.TimeoutExpired as ex:  # a hang is a failure with the workers' progress marks, not a silent kill
        def _s(b):
            return b.decode(errors="replace") if isinstance(b, bytes) else (b or "")
        pytest.fail("2-GPU worker timed out; stdout tail: " + _s(ex.stdout)[-3000:] + " stderr tail: " + _s(ex.stderr)[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2


# ---- single-process multi-GPU driver (papr_multi_*; what `PAPR_B200_DEVICES=N papr` runs) --------------
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("nshards", [2, 3])
def test_virtual_shards_single_process_all_fixtures(built, manifest, nshards):
    """N shards on ONE device (devices=[0]*N): range-order merge, shard-boundary tails, the sequential
    sum chained across shards - stdout identical to the reference on every fixture."""
    import struct
    m = built.MultiEngine([0] * nshards)
    m.set("chunk_bytes", 1 << 20)  # 131072-sample chunks so that even the small fixtures really split
    assert m.exchange.startswith("host")
    try:
        for name in fixtures.FIXTURES:
            img = fixtures.image(name)
            if fixtures.md5(img) != manifest[name]["input_md5"]:
                continue
            for graph in (False, True):
                want = open(os.path.join(GOLD, name + (".g.out" if graph else ".out")), "rb").read()
                res = m.analyze_host(img, graph=graph)
                assert built.format_result(res) == want, (name, graph)
        from test_oracle import edge_images
        for k, nf, img in edge_images():  # lengths on / just past a chunk (= shard) boundary, ragged bytes
            for graph in (False, True):
                assert built.format_result(m.analyze_host(img, graph=graph)) == oracle_binding.run_image(img, graph), (nf, graph)
        f = fixtures.siggen(0, 900_001, 31)  # sum chained across 3 shards == the oracle's sequential sum
        st, *_ = oracle_binding.analyze(f, False)
        res = m.analyze_host(f)
        assert struct.pack("<d", res.stats.sum) == struct.pack("<d", st.sum)
    finally:
        m.close()


def test_cli_two_gpus(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    f = fixtures.siggen(0, (1 << 24) + 4097, 41)
    p = tmp_path / "cap.cfile"
    p.write_bytes(f.tobytes()[:-4])  # odd number of floats: the lone-I tail lands in the last shard
    for graph, exchange in ((False, "host"), (True, "host"), (False, "nccl")):
        want = oracle_binding.run_image(p.read_bytes(), graph)
        env = dict(os.environ, PAPR_B200_DEVICES="2", PAPR_B200_STATS="1", PAPR_B200_CHUNK_MB="16",
                   PAPR_B200_EXCHANGE=exchange)
        r = subprocess.run([built.cli_path()] + (["-g"] if graph else []) + [str(p)], capture_output=True, env=env,
                           timeout=600)
        assert r.returncode == 0 and r.stdout == want, r.stderr.decode()
        assert ("exchange=%s" % exchange).encode() in r.stderr, r.stderr
