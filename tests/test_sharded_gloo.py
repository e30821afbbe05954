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
