"""The oracle (oracle/papr_oracle.c, our CPU restatement) pinned against the reference's own
outputs: tests/golden/*.out were produced by the unmodified reference binary built from
/root/reference/papr.c (tests/golden/make_golden.py).  CPU-only."""
import os

import numpy as np
import pytest

import fixtures
import oracle_binding

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name, graph):
    with open(os.path.join(GOLD, name + (".g.out" if graph else ".out")), "rb") as f:
        return f.read()


@pytest.mark.parametrize("graph", [False, True], ids=["1dB", "graph"])
@pytest.mark.parametrize("name", list(fixtures.FIXTURES))
def test_oracle_matches_reference_stdout(name, graph, manifest):
    img = fixtures.image(name)
    if fixtures.md5(img) != manifest[name]["input_md5"]:
        pytest.skip("fixture generator drifted from the recorded input (numpy RNG change)")
    assert oracle_binding.run_image(img, graph) == _gold(name, graph)


def test_appendix_a_known_answer(manifest):
    # language-independent golden vector
    img = fixtures.image("appA_1M")
    assert fixtures.md5(img) == "701a0ece11d6b9271d62debaee28aaf2"
    assert manifest["appA_1M"]["stdout_md5.out"] == "6ceb7c9fec5b4710f57a72280ea7e060"
    assert manifest["appA_1M"]["stdout_md5.g.out"] == "b5cbf837e98c6d348586806a4bf59ab7"
    f = np.frombuffer(img, np.float32)
    assert [float.hex(float(x)) for x in f[:4]] == [
        "0x1.0cf8000000000p-5", "0x1.9bb3000000000p-3", "-0x1.4338000000000p-5", "0x1.72ec000000000p-4"]


def test_siggen_c_twin_equals_numpy_twin():
    for first, n, seed in [(0, 4096, 1), (2 ** 32 - 100, 300, 2), (2 ** 35 + 7, 1000, 3)]:
        assert np.array_equal(oracle_binding.siggen(first, n, seed), fixtures.siggen(first, n, seed))


def test_reference_binary_if_present(manifest, tmp_path):
    """Where oracle/_ref/papr exists (it travels with the snapshot) re-run it on a few fixtures."""
    import subprocess
    ref = os.path.join(oracle_binding.ORACLE_DIR, "_ref", "papr")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/papr not built")
    for name in ("appA_300k_s7", "odd_bytes_odd", "burst"):
        p = tmp_path / (name + ".cfile")
        p.write_bytes(fixtures.image(name))
        for graph in (False, True):
            r = subprocess.run([ref] + (["-g"] if graph else []) + [str(p)], capture_output=True)
            assert r.stdout == _gold(name, graph)


def edge_images():
    """Captures whose length sits on / just past a multiple of 131072 samples (the 1 MiB chunks the GPU
    tests force), with 0-3 ragged trailing bytes; shared with tests/test_gpu_parity.py."""
    cs = 131072
    base = fixtures.siggen(0, 6 * cs + 10, 77)
    for k, nf in enumerate([2 * cs, 2 * cs + 1, 4 * cs + 1, 2 * (5 * cs) + 1, 2 * (6 * cs) + 3, 2 * cs - 1, 1, 2]):
        yield k, nf, base[:nf].tobytes() + bytes([1, 2, 3][: k % 4])


def test_oracle_equals_reference_binary_on_chunk_boundary_captures(tmp_path):
    """The GPU edge-case test checks against the oracle; this pins the oracle itself to the unmodified
    reference binary on the same captures (lone-I tails paired with stale Q, ragged bytes)."""
    import subprocess
    ref = os.path.join(oracle_binding.ORACLE_DIR, "_ref", "papr")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/papr not built")
    for k, nf, img in edge_images():
        p = tmp_path / ("edge%d.cfile" % k)
        p.write_bytes(img)
        for graph in (False, True):
            r = subprocess.run([ref] + (["-g"] if graph else []) + [str(p)], capture_output=True)
            assert r.stdout == oracle_binding.run_image(img, graph), (k, nf, graph)


def random_captures(seed=20261017, count=48):
    """Seeded small captures of very different character: scales from denormal to 1e15, heavy tails, constant and
    all-zero stretches, signed zeros, NaN / Inf samples, ragged byte lengths."""
    rng = np.random.default_rng(seed)
    for k in range(count):
        n = int(rng.integers(1, 6000))
        kind = k % 8
        if kind == 0:
            f = rng.standard_normal(2 * n) * 10.0 ** rng.uniform(-22, 15)
        elif kind == 1:
            f = rng.uniform(-1, 1, 2 * n) * 10.0 ** rng.uniform(-3, 3)
        elif kind == 2:  # heavy tail: a few huge samples among small ones (large PAPR, many levels)
            f = rng.standard_normal(2 * n) * 1e-3
            f[rng.integers(0, 2 * n, 3)] = rng.choice([-1.0, 1.0], 3) * 10.0 ** rng.uniform(0, 6, 3)
        elif kind == 3:  # constant envelope / exact ties between extrema
            f = np.tile(np.array([0.5, -0.5], np.float64) * rng.choice([1.0, 3.0, 0.1]), n)
        elif kind == 4:  # zeros, then signal, signed zeros inside
            f = rng.standard_normal(2 * n)
            f[: 2 * (n // 2)] = 0.0
            f[rng.integers(0, 2 * n, 4)] = -0.0
        elif kind == 5:  # denormal powers
            f = rng.standard_normal(2 * n) * 1e-21
        elif kind == 6:  # a NaN or an Inf somewhere (the reference's prints become nan / inf)
            f = rng.standard_normal(2 * n)
            f[int(rng.integers(0, 2 * n))] = rng.choice([np.nan, -np.nan, np.inf, -np.inf])
        else:  # small integers: exact powers, many equal samples
            f = rng.integers(-4, 5, 2 * n).astype(np.float64)
        img = f.astype(np.float32).tobytes()
        yield k, img[: len(img) - int(rng.integers(0, 8))]  # 0-7 bytes short: lone I, ragged tails


def test_oracle_equals_reference_binary_on_random_captures(tmp_path):
    """Differential test of the restatement against the unmodified reference binary (oracle/_ref/papr) on 48
    seeded captures x 2 modes - beyond the committed golden vectors."""
    import subprocess
    ref = os.path.join(oracle_binding.ORACLE_DIR, "_ref", "papr")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/papr not built")
    for k, img in random_captures():
        p = tmp_path / ("rnd%d.cfile" % k)
        p.write_bytes(img)
        for graph in (False, True):
            r = subprocess.run([ref] + (["-g"] if graph else []) + [str(p)], capture_output=True, timeout=120)
            assert r.stdout == oracle_binding.run_image(img, graph), (k, len(img), graph)
