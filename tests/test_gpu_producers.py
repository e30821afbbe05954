"""GPU tests of the §8(f) neighbours of the hot path: the device-side IQ producers and the
tone-reservation PAPR-reduction kernel (§8f-4), each evaluated with the CCDF engine itself."""
import os
import sys

import numpy as np
import pytest

import fixtures
import oracle_binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(built):
    e = built.Engine(0)
    yield e
    e.close()


def test_ofdm_producer_pinned_to_the_numpy_fixture(built, eng):
    """ofdm_capture with the cells of tests/fixtures._ofdm32k (numpy, float64 IFFT, the `ofdm32k` golden fixture):
    same capture to float32 rounding, and the SAME stdout from the engine as the golden file of the reference."""
    import torch
    from dtv_utils_b200.producers import ofdm_capture
    nsym, nact = 6, 27265
    rng = np.random.default_rng(32768)
    idx = np.empty((nsym, nact, 2), np.int64)
    for s in range(nsym):  # the fixture draws I then Q per symbol
        idx[s, :, 0] = rng.integers(0, 16, nact)
        idx[s, :, 1] = rng.integers(0, 16, nact)
    want = np.frombuffer(fixtures.image("ofdm32k"), np.float32)
    n = want.size // 2
    got = ofdm_capture(n, device="cuda:0", base_symbols=nsym, indices=torch.from_numpy(idx)).cpu().numpy()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(want))  # float32 cuFFT vs float64 numpy
    p = float(np.mean(got.astype(np.float64) ** 2) * 2)
    assert abs(p - 0.04) < 1e-3  # 0.2^2 (dvbt2-blade.py:132)


@pytest.mark.parametrize("which", ["dvbt2_p1", "dvbt_8k", "dvbt_2k"])
def test_producer_tails_against_oracle(built, eng, which):
    """DVB-T2 with the P1 preamble (dvbt2-blade.py:131) and the DVB-T tail (dvbt-blade.py:187-189): mean power as
    the chains' constants imply, and the engine's text on the produced capture == the oracle's."""
    from dtv_utils_b200.producers import ofdm_capture, dvbt_capture
    n = (1 << 21) + 777
    if which == "dvbt2_p1":
        d = ofdm_capture(n, seed=3, device="cuda:0", base_symbols=8, p1=True)
        power = 0.04
    else:
        d = dvbt_capture(n, seed=4, device="cuda:0", mode=which[-2:], base_symbols=16)
        nfft, nact = (2048, 1705) if which.endswith("2k") else (8192, 6817)
        power = nact * 0.0022097087 ** 2  # N*ifft of unit-power cells has mean power nact
    host = d.cpu().numpy()
    p = float(np.mean(host.astype(np.float64) ** 2) * 2)
    assert abs(p - power) < 0.03 * power, (p, power)
    for graph in (False, True):
        want = oracle_binding.run_image(host.tobytes(), graph)
        assert built.format_result(eng.analyze_device(d, n, graph)) == want


@pytest.mark.parametrize("fft_size,active,ntones", [(4096, 3409, 36), (32768, 27265, 288)])
def test_tone_reservation_kernel_against_numpy_restatement(built, eng, fft_size, active, ntones):
    """papr_tr_reduce_device vs oracle/paprtr_oracle.py (EN 302 755 clause 9.6.2.1 in float64; parity unpinned: the
    GNU Radio block is not in the reference tree).  Tolerance: 2e-5 of the largest sample (float32 kernel, float32
    cuFFT kernel vector against float64), iteration counts identical."""
    import torch
    import paprtr_oracle
    from dtv_utils_b200.producers import ofdm_symbols, tr_tones, tr_kernel
    nsym = 5
    tones = tr_tones(fft_size, active, ntones)
    x = ofdm_symbols(nsym, seed=9, device="cuda:0", fft_size=fft_size, active=active, reserved=tones)
    x0 = x.cpu().numpy().astype(np.complex128)
    p = tr_kernel(fft_size, tones, "cuda:0")
    for vclip, amax in ((3.3, 1e30), (2.9, 0.05)):  # the second: the per-tone bound Amax cuts steps short
        xs = x.clone()
        r, it = eng.tr_reduce(xs, p, tones, vclip, 3, amax)
        want_x, want_r, want_it = paprtr_oracle.tr_reduce(x0, paprtr_oracle.tr_kernel(fft_size, tones), tones, vclip, 3, amax)
        assert it.cpu().numpy().tolist() == want_it.tolist()
        scale = np.max(np.abs(x0))
        assert np.max(np.abs(xs.cpu().numpy() - want_x)) <= 2e-5 * scale
        assert np.max(np.abs(r.cpu().numpy() - want_r)) <= 1e-4 * max(1e-30, np.max(np.abs(want_r)))
        if amax > 1:  # peaks really came down to (about) the clipping level wherever the iterations sufficed
            peak = np.max(np.abs(xs.cpu().numpy()), axis=1)
            assert np.all(peak <= np.max(np.abs(x0), axis=1) + 1e-6)
    # only the reserved tones changed in the frequency domain
    xs = x.clone()
    eng.tr_reduce(xs, p, tones, 3.0, 3, 1e30)
    diff = torch.fft.fft(xs - x, dim=1).abs().cpu().numpy()
    mask = np.ones(fft_size, bool)
    mask[tones] = False
    assert diff[:, mask].max() <= 1e-3 * diff.max()


def test_tone_reservation_lowers_the_ccdf(built, eng):
    """The natural consumer of a fast CCDF: the same cells with and without tone reservation,
    both analysed by the engine - the PAPR and the high-level tail of the CCDF come down."""
    from dtv_utils_b200.producers import ofdm_capture
    n = 33024 * 96
    plain = ofdm_capture(n, seed=21, device="cuda:0", base_symbols=96)
    tr = ofdm_capture(n, seed=21, device="cuda:0", base_symbols=96, engine=eng, tone_reservation=True)
    a, b = eng.analyze_device(plain, n, False), eng.analyze_device(tr, n, False)
    assert built.format_result(b) == oracle_binding.run_image(tr.cpu().numpy().tobytes(), False)
    assert b.papr < a.papr - 0.5, (a.papr, b.papr)
    # 3.3 x (unit mean power) x 0.2: the clipping level; the tail above it must have thinned out
    lvl = (3.3 * 0.2) ** 2
    ca = int((np.frombuffer(plain.cpu().numpy().tobytes(), np.float32).reshape(-1, 2) ** 2).sum(1).__gt__(lvl).sum())
    cb = int((np.frombuffer(tr.cpu().numpy().tobytes(), np.float32).reshape(-1, 2) ** 2).sum(1).__gt__(lvl).sum())
    assert cb < ca
