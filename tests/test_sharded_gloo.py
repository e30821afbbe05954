"""CPU-only, world_size 2, gloo: the byte-range sharding logic of dtv_utils_b200.analyze_sharded
(all-gather of pass-1 states merged in rank order + all-reduce of the level counts).  The per-shard
compute is supplied by a stand-in engine built on the oracle (test infrastructure), so what is under
test is the host-side orchestration, index offsets, tie-breaking across ranks and the fused-miss
fallback protocol — not the kernels (tests/test_gpu_parity.py covers those through the C ABI)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, graph, fused, exact, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import dtv_utils_b200 as pb
    import fixtures
    import oracle_binding
    from test_abi_host import _stats_from_oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class OracleShardEngine:
        """same duck type as pb.Engine's shard stages"""

        def stats_shard(self, iq, n, first):
            st, *_ = oracle_binding.analyze(iq[:2 * n], False)
            s = _stats_from_oracle(pb, st)
            for f, v in (("peak_idx", s.peak), ("re_pos_idx", s.re_pos), ("im_pos_idx", s.im_pos),
                         ("re_neg_idx", s.re_neg), ("im_neg_idx", s.im_neg)):
                if v != 0.0:
                    setattr(s, f, getattr(s, f) + first)
            return s

        def ccdf_shard(self, iq, n, level):
            lv = np.array(level, np.float32)
            cnt = np.zeros(max(len(lv), 1), np.int64)
            oracle_binding.load().papr_oracle_pass2(iq.ctypes.data, n, lv.ctypes.data, len(lv), cnt.ctypes.data)
            return cnt[:len(lv)].tolist()

        # fused protocol: rank 1 always reports a miss -> every rank must fall back
        def fused_presample(self, iq, n, graph=False):
            return [1.0, 2.0, float(n), 0.0]

        def fused_scan(self, iq, n, first, pre, graph):
            assert pre[2] == float(self.total)  # pre[] was summed over ranks
            return self.stats_shard(iq, n, first)

        def fused_counts(self, merged, graph):
            return rank == 1, [0] * 2048

        # exact sequential sum stages: literal float64 adds in file order (np.cumsum is sequential)
        def _powers(self, iq, n):
            i, q = iq[0:2 * n:2], iq[1:2 * n:2]
            return (i * i + q * q).astype(np.float32).astype(np.float64)  # three float32 roundings, papr.c:103

        def seqsum_prepare(self, iq, n):
            return float(np.sum(self._powers(iq, n)))

        def seqsum_runs(self, iq, n, pre):
            self.pre_seen = pre
            return True

        def seqsum_chain(self, iq, n, state):
            return float(np.cumsum(np.concatenate([[state], self._powers(iq, n)]))[-1])

    f = np.frombuffer(fixtures.image(name), np.float32)
    ntot = f.size // 2
    cut = (ntot // 2) & ~1
    lo, hi = (0, cut) if rank == 0 else (cut, ntot)
    eng = OracleShardEngine()
    eng.total = ntot
    shard = np.ascontiguousarray(f[2 * lo:2 * hi])
    res = pb.analyze_sharded(eng, shard, hi - lo, lo, graph,
                             mode=pb.papr.MODE_FUSED if fused else pb.papr.MODE_TWO_PASS, exact_sum=exact)
    if exact:  # the running sum was handed from rank 0 to rank 1: equal to the whole-file sequential sum
        import struct
        whole, *_ = oracle_binding.analyze(f, False)
        assert struct.pack("<d", res.stats.sum) == struct.pack("<d", whole.sum)
        if rank == 1:
            assert abs(eng.pre_seen - float(np.sum(eng._powers(f, cut)))) < 1e-9 * max(1e-300, abs(eng.pre_seen))
    q.put((rank, pb.format_result(res)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,graph,fused,exact", [("ties", False, False, False), ("appA_300k_s7", True, False, True),
                                                     ("burst", False, True, False), ("neg_only", True, True, True),
                                                     ("gauss_1M", False, False, True)])
def test_two_rank_sharding_matches_reference(name, graph, fused, exact, built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + hash((name, graph)) % 97) % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, graph, fused, exact, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = open(os.path.join(ROOT, "tests", "golden", name + (".g.out" if graph else ".out")), "rb").read()
    # only fixtures with an even float count are used here, so no stale-Q tail is involved
    assert outs[0] == want and outs[1] == want


def _attach_worker(rank, world, port, fail_rank, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import dtv_utils_b200 as pb

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Windows:
        """stand-in for the engine's exchange-window calls (the real ones need a GPU)"""
        attached = detached = False

        def xchg_export(self):
            if rank == fail_rank == 0:
                raise pb.PaprError("no peer access")
            return bytes([rank]) * 64

        def xchg_attach(self, r, w, handles):
            assert (r, w) == (rank, world) and len(handles) == 64 * world
            assert all(handles[64 * k] == k for k in range(world))  # gathered in rank order
            if rank == fail_rank == 1:
                raise pb.PaprError("cudaIpcOpenMemHandle failed")
            self.attached = True

        def xchg_detach(self):
            self.detached = True

    eng = Windows()
    ok = pb.attach_peer_exchange(eng)
    q.put((rank, ok, eng.attached, eng.detached))
    eng._p2p = ok
    pb.detach_peer_exchange(eng)
    assert eng._p2p is None
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 0, 1], ids=["all_attach", "export_fails_on_0", "attach_fails_on_1"])
def test_peer_exchange_attach_is_all_or_nothing(fail_rank, built):
    """attach_peer_exchange: handles gathered in rank order; if ANY rank cannot export or map, EVERY
    rank reports False (and a rank that did attach lets go again), so all ranks pick the same path."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + fail_rank) % 90
    procs = [ctx.Process(target=_attach_worker, args=(r, 2, port, fail_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = {r: (ok, att, det) for r, ok, att, det in (q.get(timeout=120) for _ in procs)}
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    if fail_rank < 0:
        assert outs == {0: (True, True, False), 1: (True, True, False)}
    else:
        assert outs[0][0] is False and outs[1][0] is False
        assert all(det or not att for _, att, det in outs.values())  # whoever attached has detached again
