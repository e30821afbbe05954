"""CPU-only: the C-ABI library loads and exports what include/papr_b200.h declares; the host-side
stages (levels / merge / format / CLI argument handling) reproduce the reference's text.  No GPU
compute is called here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import fixtures
import oracle_binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(built):
    lib = built.load_library()
    hdr = open(os.path.join(ROOT, "include", "papr_b200.h")).read()
    declared = set(re.findall(r"\b(papr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(built.papr.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.papr_abi_version() == 1


def test_struct_layouts_match_header(built, tmp_path):
    """ctypes mirrors == what a C compiler lays out from include/papr_b200.h."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "papr_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(papr_stats), sizeof(papr_result),'
                   ' offsetof(papr_result, level), offsetof(papr_result, level_count), offsetof(papr_result, mode_used),'
                   ' offsetof(papr_result, h2d_bytes), offsetof(papr_stats, peak_idx));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    R, S = built.PaprResult, built.PaprStats
    assert got == [C.sizeof(S), C.sizeof(R), R.level.offset, R.level_count.offset, R.mode_used.offset,
                   R.h2d_bytes.offset, S.peak_idx.offset]


def _stats_from_oracle(pb, st):
    s = pb.PaprStats()
    s.n, s.sum, s.peak, s.peak_idx = st.n, st.sum, st.peak, st.peak_idx
    s.re_pos, s.im_pos, s.re_neg, s.im_neg = st.re_pos, st.im_pos, st.re_neg, st.im_neg
    s.re_pos_idx, s.im_pos_idx, s.re_neg_idx, s.im_neg_idx = (st.re_pos_idx, st.im_pos_idx,
                                                                st.re_neg_idx, st.im_neg_idx)
    return s


@pytest.mark.parametrize("graph", [False, True], ids=["1dB", "graph"])
@pytest.mark.parametrize("name", ["appA_1M", "gauss_1M", "const_4k", "denorm", "burst", "zeros_1000",
                                  "ties", "neg_only", "ofdm32k", "one_sample"])
def test_host_epilogue_and_text(built, name, graph, manifest):
    """papr_levels + papr_format on oracle statistics/counts == the reference's stdout."""
    img = fixtures.image(name)
    if fixtures.md5(img) != manifest[name]["input_md5"]:
        pytest.skip("fixture drift")
    iq = np.frombuffer(img, np.float32)
    st, avg, papr, level, counts = oracle_binding.analyze(iq, graph)
    s = _stats_from_oracle(built, st)
    a2, p2, lv2 = built.levels(s, graph)
    assert (a2 == avg or (a2 != a2 and avg != avg)) and (p2 == papr or (p2 != p2 and papr != papr))
    assert np.array_equal(np.array(lv2, np.float32), level)
    res = built.papr.result_from_parts(s, graph, counts)
    want = open(os.path.join(GOLD, name + (".g.out" if graph else ".out")), "rb").read()
    assert built.format_result(res) == want


def test_merge_is_first_occurrence(built):
    iq = np.frombuffer(fixtures.image("ties"), np.float32)
    whole, *_ = oracle_binding.analyze(iq, False)
    cuts = [0, 100, 101, 400, 1000, 4096]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        st, *_ = oracle_binding.analyze(iq[2 * a:2 * b], False)
        s = _stats_from_oracle(built, st)
        for f in ("peak_idx", "re_pos_idx", "im_pos_idx", "re_neg_idx", "im_neg_idx"):
            # oracle indices are range-local; "no occurrence" stays 0 like the reference's init
            val = {"peak_idx": s.peak, "re_pos_idx": s.re_pos, "im_pos_idx": s.im_pos,
                   "re_neg_idx": s.re_neg, "im_neg_idx": s.im_neg}[f]
            if val != 0.0:
                setattr(s, f, getattr(s, f) + a)
        parts.append(s)
    m = built.merge_stats(parts)
    w = _stats_from_oracle(built, whole)
    assert m.as_tuple()[2:] == w.as_tuple()[2:] and m.n == w.n
    assert abs(m.sum - w.sum) <= 1e-12 * abs(w.sum)


def _cli(built, *args):
    return subprocess.run([built.cli_path(), *args], capture_output=True)


def test_cli_usage_and_open_errors_match_reference(built, tmp_path):
    """papr.c:53-58,84-98: usage text, exit(-1) == 255, open failure message — no GPU involved."""
    usage = b"usage: papr -g <infile>\nOptions:\n\tg = graph suitable output\n"
    for args in ([], ["a", "b", "c"], ["x", "y"]):
        r = _cli(built, *args)
        assert (r.returncode, r.stdout, r.stderr) == (255, b"", usage)
    r = _cli(built, "/nonexistent/file.cfile")
    assert (r.returncode, r.stderr) == (255, b"Cannot open bitstream file </nonexistent/file.cfile>\n")
    r = _cli(built, "-gx", "/nonexistent/f")
    assert r.returncode == 255
    assert r.stderr == b"Unsupported Option: x\nCannot open bitstream file </nonexistent/f>\n"
    ref = os.path.join(ROOT, "oracle", "_ref", "papr")
    if os.path.exists(ref):
        for args in ([], ["x", "y"], ["/nonexistent/f"], ["-gx", "/nonexistent/f"]):
            a, b = _cli(built, *args), subprocess.run([ref, *args], capture_output=True)
            assert (a.returncode, a.stdout, a.stderr) == (b.returncode, b.stdout, b.stderr)


def test_product_fails_loudly_without_gpu(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = tmp_path / "x.cfile"
    p.write_bytes(fixtures.image("one_sample"))
    r = _cli(built, str(p))
    assert r.returncode == 1 and r.stdout == b"" and b"GPU engine unavailable" in r.stderr
    with pytest.raises(built.PaprError):
        built.Engine(0)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under dtv-utils_b200/ may mention it."""
    pkg = os.path.join(ROOT, "dtv-utils_b200")
    for d, _, files in os.walk(pkg):
        if os.path.basename(d) in ("build", "lib", "bin", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("the oracle", ""), os.path.join(d, f)


# ---- the tables that let the GPU evaluate the reference's epilogue (papr.c:131-141 / 164-173) ----------
def _device_epilogue(tables, summ, n, peak, L_max):
    """What papr_levels_kernel (csrc/papr_kernels.cu: merge_and_levels) computes, restated with numpy's
    IEEE double divide / multiply / compare and one rounding to float32 — no libm."""
    pow10, ratio_min = tables
    avg = np.float64(summ) / np.float64(n)
    ratio = np.float64(np.float32(peak)) / avg
    L = int(np.searchsorted(ratio_min[:L_max], ratio, side="right"))  # number of j with ratio >= ratio_min[j]
    with np.errstate(over="ignore"):
        level = (pow10[:L] * avg).astype(np.float32)
    return L, level


@pytest.mark.parametrize("graph", [False, True], ids=["1dB", "graph"])
def test_device_epilogue_tables_agree_with_host_libm(built, graph):
    """papr_host_build_tables (host libm, built once per engine) + plain IEEE arithmetic must give the
    same number of levels and bit-identical thresholds as papr_levels (the reference's conversion
    sequence with log10 / pow) — for thousands of random statistics, with peak/avg ratios placed on
    and next to the dB boundaries where (int)papr steps."""
    lib = built.load_library()
    nmax = 2048 if graph else 256
    pow10 = np.full(2048, np.inf)
    ratio_min = np.full(2048, np.inf)
    lib.papr_host_build_tables.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.papr_host_build_tables.restype = None
    lib.papr_host_build_tables(int(graph), nmax, pow10.ctypes.data, ratio_min.ctypes.data)
    assert np.all(np.diff(ratio_min[:nmax]) >= 0) and ratio_min[0] > 0
    rng = np.random.default_rng(7)
    cases = []
    for _ in range(3000):
        n = int(rng.integers(1, 1 << 40))
        avg = float(np.exp(rng.uniform(-40, 20)))
        papr_db = float(rng.uniform(0, min(60.0, 10 * np.log10(n))))  # peak <= sum: the peak is one of the samples
        if rng.random() < 0.5:  # sit on a boundary of the level count (1 dB or 0.1 dB steps) and step around it
            papr_db = round(papr_db, 0 if not graph else 1)
        peak = np.float32(avg * 10 ** (papr_db / 10))
        for nudge in (0, 1, -1):
            pk = np.nextafter(peak, np.float32(np.inf if nudge > 0 else -np.inf)) if nudge else peak
            cases.append((avg * n, n, float(pk)))
    cases += [(1.0, 1, 1.0), (4.0, 4, 1.0), (4.0, 4, 4.0), (2.0 ** 60, 2 ** 62, 2.0 ** 60), (1e-40, 1000, 1e-42)]
    st = built.PaprStats()
    lv = (C.c_float * 2048)()
    avg_c, papr_c = C.c_double(), C.c_float()
    for summ, n, peak in cases:
        st.n, st.sum, st.peak = n, summ, peak
        L_host = lib.papr_levels(C.byref(st), int(graph), C.byref(avg_c), C.byref(papr_c), lv, 2048)
        L_dev, level_dev = _device_epilogue((pow10, ratio_min), st.sum, n, st.peak, nmax)
        assert L_dev == L_host, (summ, n, peak, papr_c.value)
        host = np.frombuffer(lv, np.float32, L_host)
        assert host.tobytes() == level_dev.tobytes(), (summ, n, peak)
