"""GPU parity tests (run by the driver with -m gpu on a B200): the CUDA path, called through the
C ABI (ctypes -> libpapr_b200.so) and through the drop-in CLI, against
  * the committed golden stdout of the unmodified reference binary (tests/golden/), bit for bit;
  * the oracle restatement on seeded inputs the oracle finishes in seconds;
  * size-independent properties at BASELINE.json's full single-GPU size.
Tolerance: NONE — integer counts, offsets and every printed character must be identical.  (The one
floating-point accumulator, the double sum, is checked to 1e-12 relative where it is compared
directly; see DESIGN.md "the sequential sum".)"""
import os
import subprocess

import numpy as np
import pytest

import fixtures
import oracle_binding

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _gold(name, graph):
    return open(os.path.join(GOLD, name + (".g.out" if graph else ".out")), "rb").read()


@pytest.fixture(scope="module")
def eng(built):
    e = built.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


def _dev(torch, f32):
    return torch.from_numpy(np.ascontiguousarray(f32)).to("cuda:0")


# ---- golden vectors of the reference binary --------------------------------------------------------
@pytest.mark.parametrize("graph", [False, True], ids=["1dB", "graph"])
@pytest.mark.parametrize("name", list(fixtures.FIXTURES))
def test_cli_stdout_equals_reference(built, name, graph, manifest, tmp_path):
    img = fixtures.image(name)
    if fixtures.md5(img) != manifest[name]["input_md5"]:
        pytest.skip("fixture drift")
    p = tmp_path / (name + ".cfile")
    p.write_bytes(img)
    r = subprocess.run([built.cli_path()] + (["-g"] if graph else []) + [str(p)], capture_output=True)
    assert r.returncode == 0 and r.stderr == b""
    assert r.stdout == _gold(name, graph)


@pytest.mark.parametrize("graph", [False, True], ids=["1dB", "graph"])
@pytest.mark.parametrize("name", list(fixtures.FIXTURES))
def test_host_buffer_path_equals_reference(built, eng, name, graph, manifest):
    img = fixtures.image(name)
    if fixtures.md5(img) != manifest[name]["input_md5"]:
        pytest.skip("fixture drift")
    for chunk in (64 << 20, 1 << 20):  # one chunk / many chunks (+ the staging ring wrapping around)
        eng.set("chunk_bytes", chunk)
        assert built.format_result(eng.analyze_host(img, graph=graph)) == _gold(name, graph)
    eng.set("chunk_bytes", 64 << 20)


@pytest.mark.parametrize("mode", [1, 2], ids=["two_pass", "fused"])
@pytest.mark.parametrize("graph", [False, True], ids=["1dB", "graph"])
@pytest.mark.parametrize("name", [n for n in fixtures.FIXTURES if not n.startswith("odd")])
def test_device_path_equals_reference(built, eng, torch_cuda, name, graph, mode, manifest):
    img = fixtures.image(name)
    if fixtures.md5(img) != manifest[name]["input_md5"]:
        pytest.skip("fixture drift")
    f = np.frombuffer(img, np.float32)
    d = _dev(torch_cuda, f) if f.size else torch_cuda.empty(4, device="cuda:0")
    eng.set("mode", mode)
    try:
        res = eng.analyze_device(d, f.size // 2, graph)
    finally:
        eng.set("mode", 0)
    assert built.format_result(res) == _gold(name, graph)
    assert res.mode_used == mode


@pytest.mark.parametrize("graph", [False, True], ids=["1dB", "graph"])
@pytest.mark.parametrize("name", list(fixtures.FIXTURES))
def test_capture_larger_than_hbm_is_restreamed(built, eng, name, graph, manifest, tmp_path):
    """A capture that does not fit the resident budget (forced here: 1 MiB) is streamed twice through a
    ring of device chunk buffers - the reference reads its file twice as well (papr.c:142).  Host image
    and file (pread) sources, all fixtures incl. NaN sign, lone-I tails and the empty file."""
    img = fixtures.image(name)
    if fixtures.md5(img) != manifest[name]["input_md5"]:
        pytest.skip("fixture drift (different numpy RNG stream)")
    path = tmp_path / "capture.cfile"
    path.write_bytes(img)
    eng.set("chunk_bytes", 1 << 20)
    eng.set("max_resident_bytes", 1 << 20)
    try:
        assert built.format_result(eng.analyze_host(img, graph=graph)) == _gold(name, graph)
        assert built.format_result(eng.analyze_file(str(path), graph)) == _gold(name, graph)
    finally:
        eng.set("max_resident_bytes", 0)
        eng.set("chunk_bytes", 64 << 20)


def test_chunk_boundary_cases_all_sources(built, eng, torch_cuda, tmp_path):
    """Captures whose length sits exactly on / just past a chunk boundary (1 MiB chunks = 131072 samples):
    a last chunk that holds nothing but the lone-I tail sample, ragged trailing bytes, more chunks than
    staging slots and ring slots - through every source (pageable image, pinned image, file) and both
    residency modes, against the oracle."""
    from test_oracle import edge_images
    eng.set("chunk_bytes", 1 << 20)
    eng.set("staging_threads", 2)
    try:
        for k, nf, img in edge_images():
            path = tmp_path / ("edge%d.cfile" % k)
            path.write_bytes(img)
            pinned = torch_cuda.from_numpy(np.frombuffer(img[: len(img) // 4 * 4], np.float32).copy()).pin_memory()
            for graph in (False, True):
                want = oracle_binding.run_image(img, graph)
                for budget in (0, 1 << 20):
                    eng.set("max_resident_bytes", budget)
                    assert built.format_result(eng.analyze_host(img, graph=graph)) == want, (nf, graph, budget, "image")
                    assert built.format_result(eng.analyze_file(str(path), graph)) == want, (nf, graph, budget, "file")
                    if len(img) % 4 == 0:  # a pinned tensor cannot carry the ragged bytes
                        assert built.format_result(eng.analyze_host(pinned, graph=graph)) == want, (nf, graph, budget, "pinned")
    finally:
        eng.set("max_resident_bytes", 0)
        eng.set("staging_threads", -1)
        eng.set("chunk_bytes", 64 << 20)


def test_random_captures_host_and_device_paths(built, eng, torch_cuda):
    """The 48 seeded captures of tests/test_oracle.py (pinned there to the reference binary: scales from denormal to
    1e15, heavy tails, constant stretches, signed zeros, NaN / Inf, ragged lengths) through the host path and,
    where the bytes are whole 16-byte-aligned samples, the device path in both schedules."""
    from test_oracle import random_captures
    try:
        for k, img in random_captures():
            for graph in (False, True):
                want = oracle_binding.run_image(img, graph)
                assert built.format_result(eng.analyze_host(img, graph=graph)) == want, (k, len(img), graph, "host")
                if len(img) % 8 == 0 and len(img) >= 8:
                    d = _dev(torch_cuda, np.frombuffer(img, np.float32))
                    for mode in (1, 2):
                        eng.set("mode", mode)
                        got = built.format_result(eng.analyze_device(d, len(img) // 8, graph))
                        assert got == want, (k, len(img), graph, mode)
                    eng.set("mode", 0)
    finally:
        eng.set("mode", 0)


def test_pinned_source_goes_direct(built, eng, torch_cuda):
    f = torch_cuda.from_numpy(fixtures.siggen(0, 300_000, 11)).pin_memory()
    res = eng.analyze_host(f, graph=True)
    assert built.format_result(res) == oracle_binding.run_image(f.numpy().tobytes(), True)
    assert 300_000 * 8 <= res.h2d_bytes < 300_000 * 8 + 4096  # the capture once, plus a few control bytes


# ---- seeded inputs against the oracle --------------------------------------------------------------
def test_device_generator_equals_c_twin(eng, torch_cuda):
    for first, n, seed in [(0, 100_003, 1), (2 ** 32 - 1000, 5000, 2), (2 ** 35 + 12345, 70_001, 3)]:
        d = torch_cuda.empty(2 * n, dtype=torch_cuda.float32, device="cuda:0")
        eng.siggen(d, first, n, seed)
        assert np.array_equal(d.cpu().numpy(), oracle_binding.siggen(first, n, seed))


@pytest.mark.parametrize("nsamples,seed", [(1 << 22, 1), ((1 << 24) + 12345, 2)])
def test_seeded_against_oracle_all_modes(built, eng, torch_cuda, nsamples, seed):
    d = torch_cuda.empty(2 * nsamples, dtype=torch_cuda.float32, device="cuda:0")
    eng.siggen(d, 0, nsamples, seed)
    host = d.cpu().numpy()
    for graph in (False, True):
        st, avg, papr, level, counts = oracle_binding.analyze(host, graph)
        want = oracle_binding.run_image(host.tobytes(), graph)
        for mode in (1, 2):
            eng.set("mode", mode)
            res = eng.analyze_device(d, nsamples, graph)
            eng.set("mode", 0)
            assert res.counts() == counts.tolist()  # (a fused miss at these small sizes re-runs exactly)
            assert np.array_equal(np.array(res.levels(), np.float32), level)
            assert abs(res.stats.sum - st.sum) <= 1e-12 * st.sum
            assert built.format_result(res) == want
        assert built.format_result(eng.analyze_host(host, graph=graph)) == want


def test_virtual_shards_on_one_gpu(built, eng, torch_cuda):
    """N byte-range shards through the stage API + host merge == the whole capture."""
    f = np.frombuffer(fixtures.image("appA_1M"), np.float32)
    d = _dev(torch_cuda, f)
    n = f.size // 2
    for graph in (False, True):
        for nshard in (2, 3, 8):
            cuts = [((n * k // nshard) + 1) & ~1 for k in range(nshard)] + [n]  # float4-aligned cuts
            parts = [eng.stats_shard(d[2 * a:], b - a, a) for a, b in zip(cuts[:-1], cuts[1:])]
            merged = built.merge_stats(parts)
            _, _, lv = built.levels(merged, graph)
            counts = np.zeros(len(lv), np.int64)
            for a, b in zip(cuts[:-1], cuts[1:]):
                counts += np.array(eng.ccdf_shard(d[2 * a:], b - a, lv), np.int64)
            res = built.papr.result_from_parts(merged, graph, counts.tolist())
            assert built.format_result(res) == _gold("appA_1M", graph)


def test_fused_stage_api_and_forced_miss(built, eng, torch_cuda):
    f = np.frombuffer(fixtures.image("gauss_1M"), np.float32)
    d = _dev(torch_cuda, f)
    n = f.size // 2
    pre = eng.fused_presample(d, n, False)
    assert pre[2] > 16
    st = eng.fused_scan(d, n, 0, pre, False)
    miss, cnt = eng.fused_counts(st, False)
    res = built.papr.result_from_parts(st, False, [cnt[j] for j in range(2048)])
    assert not miss and built.format_result(res) == _gold("gauss_1M", False)
    # a wrong prediction (mean off by 5 %) must be detected, never silently accepted
    bad = [pre[0] * 1.05, pre[1] * 1.05 ** 2, pre[2], 0.0]
    st2 = eng.fused_scan(d, n, 0, bad, False)
    assert st2.as_tuple() == st.as_tuple()
    miss2, _ = eng.fused_counts(st2, False)
    assert miss2


def test_epilogue_levels_are_verified_and_recounted(built, eng, torch_cuda):
    """Single shard, sequential sum chained on the device: the epilogue derives avg / L / levels from the
    fixed-order sum of the CTA partials while the chain is still running (papr_xt_epilogue_kernel).  The host
    accepts its counts only if those levels are bit-identical to the levels of the exact sum; a (forced)
    disagreement must be caught and the counts taken again from the cells + fine table still on the device."""
    n = (1 << 24) + 4096 * 3 + 17
    f = fixtures.siggen(0, n, 5)
    d = _dev(torch_cuda, f)
    eng.set("mode", 2)
    try:
        for graph in (False, True):
            want = oracle_binding.run_image(f.tobytes(), graph)
            r0 = eng.analyze_device(d, n, graph)
            assert r0.fused_miss == 0 and (r0.sum_path & 0xff) == 1 and built.format_result(r0) == want
            # 3e-7: a few float32 levels move by one ulp, all stay inside their windows; 1.25: every level leaves its
            # window (the speculative counts report a miss as well) - both are settled by one more tiny kernel
            for bias in (1.0 + 3e-7, 1.25, 0.5):
                eng.set("epilogue_bias", bias)
                r = eng.analyze_device(d, n, graph)
                assert r.fused_miss == 2, (graph, bias, r.fused_miss)
                assert r.stats.as_tuple() == r0.stats.as_tuple() and r.counts() == r0.counts()
                assert built.format_result(r) == want
            eng.set("epilogue_bias", 1.0)
    finally:
        eng.set("epilogue_bias", 1.0)
        eng.set("mode", 0)


def test_generic_bsearch_kernel(built, eng, torch_cuda):
    """Force the fallback for plans whose fine table does not fit (PLAN_BSEARCH)."""
    eng.set("fine_bytes_log2", 8)
    try:
        for name in ("burst", "appA_300k_s7", "denorm"):
            f = np.frombuffer(fixtures.image(name), np.float32)
            d = _dev(torch_cuda, f)
            for graph in (False, True):
                eng.set("mode", 1)
                res = eng.analyze_device(d, f.size // 2, graph)
                assert built.format_result(res) == _gold(name, graph)
    finally:
        eng.set("fine_bytes_log2", 26)
        eng.set("mode", 0)


def test_misaligned_pointer_is_rejected(built, eng, torch_cuda):
    d = torch_cuda.zeros(1024, dtype=torch_cuda.float32, device="cuda:0")
    with pytest.raises(built.PaprError):
        eng.analyze_device(d[2:], 100, False)


# ---- properties at BASELINE.json's single-GPU size --------------------------------------------------
@pytest.mark.parametrize("log2n", [29])
def test_full_size_properties(built, eng, torch_cuda, log2n):
    """4 GiB (configs[1]): fused == two-pass; halves merge to the whole; counts are additive over
    shards, monotone in the level and bounded by N; a prefix that the oracle can afford matches."""
    n = 1 << log2n
    d = torch_cuda.empty(2 * n, dtype=torch_cuda.float32, device="cuda:0")
    eng.siggen(d, 0, n, 1)
    out = {}
    for graph in (False, True):
        for mode in (1, 2):
            eng.set("mode", mode)
            out[graph, mode] = eng.analyze_device(d, n, graph)
        eng.set("mode", 0)
        a, b = out[graph, 1], out[graph, 2]
        assert b.fused_miss == 0
        assert a.stats.as_tuple() == b.stats.as_tuple() and a.counts() == b.counts()
        assert built.format_result(a) == built.format_result(b)
        c = a.counts()
        assert all(x >= y for x, y in zip(c, c[1:])) and 0 < c[0] < n and c[-1] >= 1
    res = out[False, 1]
    h = n // 2
    parts = [eng.stats_shard(d, h, 0), eng.stats_shard(d[2 * h:], n - h, h)]
    m = built.merge_stats(parts)
    # (shard statistics carry fixed-order sums; the whole-capture result carries the reference's sequential sum)
    assert m.as_tuple()[2:] == res.stats.as_tuple()[2:] and abs(m.sum - res.stats.sum) <= 1e-10 * m.sum
    lv = res.levels()
    c2 = np.array(eng.ccdf_shard(d, h, lv)) + np.array(eng.ccdf_shard(d[2 * h:], n - h, lv))
    assert c2.tolist() == res.counts()
    # Appendix-A known answer on the first 2^20 samples of the same stream (seed 1)
    head = eng.analyze_device(d, 1 << 20, False)
    assert built.format_result(head) == _gold("appA_1M", False)


def test_two_launch_shard_32gib_fused_equals_two_pass(built, eng, torch_cuda):
    """32 GiB per GPU (BASELINE configs[2] / [3]: 2^32 samples = two launches of the sweep, tile records and the
    chain spanning both): the one-sweep analysis with the sequential sum chained on the device must print what the
    two-pass schedule prints, whose sum comes from the independent two-sweep emulation - text and sum bits."""
    import struct
    free, _total = torch_cuda.cuda.mem_get_info()
    n = 1 << 32
    if free < 8 * n + (4 << 30):
        pytest.skip("needs 36 GiB of free HBM")
    d = torch_cuda.empty(2 * n, dtype=torch_cuda.float32, device="cuda:0")
    eng.siggen(d, 0, n, 2)
    try:
        for graph in (False, True):
            eng.set("mode", 2)
            a = eng.analyze_device(d, n, graph)
            eng.set("mode", 1)
            b = eng.analyze_device(d, n, graph)
            assert (a.sum_path & 0xff) == 1 and a.fused_miss in (0, 2), (a.sum_path, a.fused_miss)
            assert (b.sum_path & 0xff) == 2
            assert struct.pack("<d", a.stats.sum) == struct.pack("<d", b.stats.sum)
            assert a.stats.as_tuple() == b.stats.as_tuple() and a.counts() == b.counts()
            assert built.format_result(a) == built.format_result(b)
            c = a.counts()
            assert all(x >= y for x, y in zip(c, c[1:])) and 0 < c[0] < n and c[-1] >= 1
    finally:
        eng.set("mode", 0)
        del d
        torch_cuda.cuda.empty_cache()


# ---- BASELINE configs[1] at full size against the REAL reference binary ------------------------------
def test_config1_4gib_cli_vs_reference_binary(built, eng, torch_cuda, tmp_path_factory):
    """`papr` on a 4 GiB synthetic capture (2^29 samples, seed 1): stdout of the drop-in CLI must be
    identical to stdout of the unmodified reference binary (oracle/_ref/papr, ~20 s of CPU)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "papr")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/papr not present")
    import shutil
    if shutil.disk_usage("/dev/shm").free < (5 << 30):
        pytest.skip("not enough tmpfs for a 4 GiB capture")
    n = 1 << 29
    path = "/dev/shm/papr_b200_test_4gib_%d.cfile" % os.getpid()
    try:
        d = torch_cuda.empty(2 * n, dtype=torch_cuda.float32, device="cuda:0")
        eng.siggen(d, 0, n, 1)
        with open(path, "wb") as f:
            step = 1 << 26
            for k in range(0, 2 * n, step):
                f.write(d[k:k + step].cpu().numpy().tobytes())
        want = subprocess.run([ref, path], capture_output=True).stdout
        got = subprocess.run([built.cli_path(), path], capture_output=True)
        assert got.returncode == 0 and got.stdout == want
        for mode in (1, 2):  # and the device-resident paths, both schedules
            eng.set("mode", mode)
            assert built.format_result(eng.analyze_device(d, n, False)) == want
        eng.set("mode", 0)
        if os.environ.get("PAPR_B200_TEST_GRAPH_4GIB"):  # ~2 min of CPU for the reference
            want_g = subprocess.run([ref, "-g", path], capture_output=True).stdout
            assert subprocess.run([built.cli_path(), "-g", path], capture_output=True).stdout == want_g
        else:  # by default -g on the first 512 MiB of the same capture (2^26 samples, ~9 s of CPU for the reference)
            m = 1 << 26
            os.truncate(path, 8 * m)
            want_g = subprocess.run([ref, "-g", path], capture_output=True).stdout
            assert subprocess.run([built.cli_path(), "-g", path], capture_output=True).stdout == want_g
            for mode in (1, 2):
                eng.set("mode", mode)
                assert built.format_result(eng.analyze_device(d, m, True)) == want_g
            eng.set("mode", 0)
    finally:
        if os.path.exists(path):
            os.unlink(path)


# ---- the sequential double sum, bit for bit (papr.c:104) -------------------------------------------
def _seq_cases():
    rng = np.random.default_rng(2024)
    n = (1 << 21) + 12345
    yield "appA", fixtures.siggen(0, n, 21)
    # 12 orders of magnitude of dynamic range: most adds round, many binade crossings
    mag = np.exp(rng.uniform(-14, 14, 2 * n)).astype(np.float32)
    yield "wide_dynamic_range", (mag * rng.choice([-1.0, 1.0], 2 * n)).astype(np.float32)
    # powers of two and small integers: exact half-ulp ties are common -> exercises round-half-even
    yield "ties", (np.ldexp(rng.integers(1, 4, 2 * n).astype(np.float32),
                            rng.integers(-14, 14, 2 * n))).astype(np.float32)
    z = fixtures.siggen(0, n, 22).copy()
    z[: 2 * 300_000] = 0.0  # long run of leading zeros, then signal
    yield "leading_zeros", z
    yield "denormal_powers", (fixtures.siggen(0, 200_000, 23) * np.float32(1e-21)).astype(np.float32)
    big = fixtures.siggen(0, 150_000, 24).copy()
    big[2 * 70_000] = 1e15  # one huge sample: the running sum jumps 30 binades mid-file
    yield "jump", big


def test_exact_sequential_sum_bit_for_bit(built, eng, torch_cuda):
    import struct
    eng.set("exact_sum", 1)
    try:
        for name, f in _seq_cases():
            f = np.ascontiguousarray(f, np.float32)
            n = f.size // 2
            st, *_ = oracle_binding.analyze(f, False)
            d = _dev(torch_cuda, f)
            for mode in (1, 2):
                eng.set("mode", mode)
                res = eng.analyze_device(d, n, False)
                assert struct.pack("<d", res.stats.sum) == struct.pack("<d", st.sum), (name, mode)
            eng.set("mode", 0)
            for graph in (False, True):  # the file/host path emulates the sequential sum by default
                res = eng.analyze_host(f, graph=graph)
                assert struct.pack("<d", res.stats.sum) == struct.pack("<d", st.sum), name
                assert built.format_result(res) == oracle_binding.run_image(f.tobytes(), graph), name
            # many chunks: tile sums / binade placement / tile runs trail the copies chunk by chunk;
            # resident shard, then the re-streaming mode (ring of device chunk buffers, two PCIe passes)
            eng.set("chunk_bytes", 1 << 20)
            eng.set("staging_threads", 2)  # 4 staging slots: they wrap around several times
            try:
                for budget in (0, 1 << 20):
                    eng.set("max_resident_bytes", budget)
                    res = eng.analyze_host(f, graph=False)
                    assert struct.pack("<d", res.stats.sum) == struct.pack("<d", st.sum), (name, budget)
                    assert built.format_result(res) == oracle_binding.run_image(f.tobytes(), False), (name, budget)
            finally:
                eng.set("max_resident_bytes", 0)
                eng.set("staging_threads", -1)
                eng.set("chunk_bytes", 64 << 20)
    finally:
        eng.set("exact_sum", -1)
        eng.set("mode", 0)


def test_exact_sequential_sum_chained_over_shards(built, torch_cuda):
    """papr_seqsum_prepare/_runs/_chain: three shards (one engine each, as three ranks would hold them),
    tile runs per shard, the running sum handed on in index order == the reference's sequential sum
    over the whole capture, bit for bit - for cuts that are NOT tile aligned."""
    import struct
    engines = [built.Engine(0) for _ in range(3)]
    try:
        for name, f in _seq_cases():
            f = np.ascontiguousarray(f, np.float32)
            n = f.size // 2
            st, *_ = oracle_binding.analyze(f, False)
            cuts = [0, (n // 3 + 1001) & ~1, (2 * n // 3 + 7) & ~1, n]
            shards = [_dev(torch_cuda, f[2 * cuts[r]:2 * cuts[r + 1]]) for r in range(3)]
            ns = [cuts[r + 1] - cuts[r] for r in range(3)]
            approx = [engines[r].seqsum_prepare(shards[r], ns[r]) for r in range(3)]
            assert abs(sum(approx) - st.sum) <= 1e-9 * abs(st.sum)
            for r in range(3):
                assert engines[r].seqsum_runs(shards[r], ns[r], sum(approx[:r])), name
            state = 0.0
            for r in range(3):
                state = engines[r].seqsum_chain(shards[r], ns[r], state)
            assert struct.pack("<d", state) == struct.pack("<d", st.sum), name
        # a non-finite sample: the emulation reports "not applicable" instead of inventing a sum
        bad = fixtures.siggen(0, 100_000, 5).copy()
        bad[12345] = np.inf
        d = _dev(torch_cuda, bad)
        engines[0].seqsum_prepare(d, 100_000)
        assert engines[0].seqsum_runs(d, 100_000, 0.0) is False
    finally:
        for e in engines:
            e.close()


def test_ofdm_producer_capture_against_oracle(built, eng, torch_cuda):
    """SURVEY §8f-1 / BASELINE configs[4]: DVB-T2-like 32K OFDM generated on the device."""
    from dtv_utils_b200.producers import ofdm_capture
    n = (1 << 22) + 33024 * 3
    d = ofdm_capture(n, seed=7, device="cuda:0")
    host = d.cpu().numpy()
    p = float(np.mean(host.astype(np.float64) ** 2) * 2)
    assert abs(p - 0.04) < 0.002  # 0.2^2: unit-power constellation through the x0.2 gain (dvbt2-blade.py:132)
    for graph in (False, True):
        want = oracle_binding.run_image(host.tobytes(), graph)
        for mode in (1, 2):
            eng.set("mode", mode)
            assert built.format_result(eng.analyze_device(d, n, graph)) == want
        eng.set("mode", 0)
        assert built.format_result(eng.analyze_host(host, graph=graph)) == want


# ---- the sequential sum INSIDE the fused sweep (papr_exact.cu): TMA-fed scan + device chain ----------
def _bits(x):
    import struct
    return struct.pack("<d", x)


@pytest.mark.parametrize("n", [8192 + 5, (1 << 17) + 255, (1 << 20) + 16 * 7 + 3, (1 << 22) + 4096 * 3 + 256 * 5 + 8,
                               (1 << 24) + 1])
def test_sum_chained_inside_the_sweep_equals_sequential_sum(built, eng, torch_cuda, n):
    """Default resident path, fused mode: stats.sum must be the reference's sequential double sum bit for
    bit AND must have been produced by the device chain (sum_path == 1), for captures whose length is not
    a multiple of the 16-sample TMA row, the 256-sample batch or the 4096-sample tile."""
    for seed, scale in ((31, 1.0), (32, 37.5), (33, 1e-9)):
        f = (fixtures.siggen(0, n, seed) * np.float32(scale)).astype(np.float32)
        st, *_ = oracle_binding.analyze(f, False)
        d = _dev(torch_cuda, f)
        eng.set("mode", 2)
        try:
            for graph in (False, True):
                res = eng.analyze_device(d, n, graph)
                assert res.sum_path & 0xff == 1, (n, seed, "device chain declined, why=%d" % (res.sum_path >> 8))
                assert _bits(res.stats.sum) == _bits(st.sum), (n, seed)
                assert built.format_result(res) == oracle_binding.run_image(f.tobytes(), graph)
        finally:
            eng.set("mode", 0)


def test_sum_chain_declines_rather_than_guessing(built, eng, torch_cuda):
    """Signals the presample cannot predict (a burst after silence, a 30-binade jump, leading zeros, a wrong
    predicted mean): the device chain must either produce the exact sum or decline (sum_path == 2) - and the
    reported sum is the sequential sum either way."""
    n = (1 << 20) + 77
    cases = {}
    z = fixtures.siggen(0, n, 41).copy()
    z[: 2 * 300_000] = 0.0
    cases["leading_zeros"] = z
    b = (fixtures.siggen(0, n, 42) * np.float32(1e-4)).astype(np.float32)
    b[2 * 700_000:] *= np.float32(3e4)
    cases["burst"] = b
    j = fixtures.siggen(0, n, 43).copy()
    j[2 * 500_000] = 1e15
    cases["jump"] = j
    eng.set("mode", 2)
    try:
        for name, f in cases.items():
            st, *_ = oracle_binding.analyze(f, False)
            res = eng.analyze_device(_dev(torch_cuda, f), n, False)
            assert res.sum_path & 0xff in (1, 2), name
            assert _bits(res.stats.sum) == _bits(st.sum), name
            assert built.format_result(res) == oracle_binding.run_image(f.tobytes(), False), name
        f = fixtures.siggen(0, n, 44)
        st, *_ = oracle_binding.analyze(f, False)
        for bias in (0.5, 1.9, 4.0):
            eng.set("predict_bias", bias)
            res = eng.analyze_device(_dev(torch_cuda, f), n, False)
            assert _bits(res.stats.sum) == _bits(st.sum), bias
            assert built.format_result(res) == oracle_binding.run_image(f.tobytes(), False), bias
    finally:
        eng.set("predict_bias", 1.0)
        eng.set("mode", 0)


# ---- non-seekable inputs: the reference needs a seekable file (papr.c:142), the drop-in does not -----------
def _feed(path_or_fd, img, piece=1 << 16):
    import threading

    def run():
        with (open(path_or_fd, "wb") if isinstance(path_or_fd, str) else os.fdopen(path_or_fd, "wb")) as f:
            for k in range(0, len(img), piece):
                f.write(img[k:k + piece])
    t = threading.Thread(target=run, daemon=True)
    t.start()
    return t


@pytest.mark.parametrize("name", ["appA_1M", "odd_floats_short", "odd_floats_long", "odd_bytes_even", "odd_bytes_odd", "odd_bytes_short",
                                  "exact_chunks", "ties", "nan", "neg_nan_q", "empty", "one_sample", "zeros_1000"])
def test_fifo_and_stdin_are_streamed(built, eng, name, manifest, tmp_path):
    """A named FIFO on the command line and `papr /dev/stdin < file` (a pipe): same text as the reference prints
    for the same bytes in a regular file - including the lone-I / stale-Q tail, which needs the last 128 KiB of
    a stream whose length is unknown until it ends.  The FIFO is opened exactly once (a writer that sees its
    reader vanish gets SIGPIPE / EPIPE)."""
    if name not in fixtures.FIXTURES:
        pytest.skip("no such fixture")
    img = fixtures.image(name)
    for graph in (False, True):
        want = _gold(name, graph)
        fifo = str(tmp_path / ("in_%d.fifo" % graph))
        os.mkfifo(fifo)
        t = _feed(fifo, img)
        r = subprocess.run([built.cli_path()] + (["-g"] if graph else []) + [fifo], capture_output=True, timeout=120)
        t.join(30)
        assert r.returncode == 0 and r.stderr == b"" and r.stdout == want, (name, graph, r.stderr[-300:])
        p = subprocess.Popen([built.cli_path()] + (["-g"] if graph else []) + ["/dev/stdin"], stdin=subprocess.PIPE,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        out, err = p.communicate(img, timeout=120)
        assert p.returncode == 0 and err == b"" and out == want, (name, graph)
    # through the C ABI with small chunks: many device chunks, the tail sample in a chunk of its own
    eng.set("chunk_bytes", 1 << 20)
    try:
        rd, wr = os.pipe()
        t = _feed(wr, img)
        res = eng.analyze_fd(rd, False)
        os.close(rd)
        t.join(30)
        assert built.format_result(res) == _gold(name, False)
    finally:
        eng.set("chunk_bytes", 64 << 20)


def test_o_direct_reads_give_the_same_answer(built, eng, manifest, tmp_path):
    """tunable "o_direct": regular files are read with O_DIRECT straight into the pinned staging slots where the
    file system allows it (silently buffered where it does not, e.g. tmpfs) - same text either way, including a
    file whose length is not a multiple of the 4096-byte alignment O_DIRECT needs."""
    eng.set("o_direct", 1)
    eng.set("chunk_bytes", 1 << 20)
    try:
        for name in ("appA_1M", "odd_bytes_odd", "exact_chunks"):
            p = tmp_path / (name + ".cfile")
            p.write_bytes(fixtures.image(name))
            for graph in (False, True):
                assert built.format_result(eng.analyze_file(str(p), graph)) == _gold(name, graph)
    finally:
        eng.set("o_direct", 0)
        eng.set("chunk_bytes", 64 << 20)
