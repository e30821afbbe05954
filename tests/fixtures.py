"""Deterministic capture fixtures shared by tests/ and tests/golden/make_golden.py.

Every fixture is a *file image* (bytes) in the reference's on-disk format: headerless host-endian
float32 I/Q pairs (reference consumer papr.c:101-103; producers dvbt2-blade.py:158-160).  The
reference ships no fixtures of its own; this set follows SURVEY.md §4.3 and pins
the edge cases listed there.  Inputs are regenerated on demand and checked against the md5 stored
in tests/golden/manifest.json, so a numpy RNG drift shows up as a skipped (not a wrong) test.
"""
from __future__ import annotations

import hashlib

import numpy as np

U64 = np.uint64


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
    return z ^ (z >> U64(31))


def siggen(first: int, nsamples: int, seed: int) -> np.ndarray:
    """SURVEY.md Appendix A generator, numpy twin (float32 array of 2*nsamples, I/Q interleaved)."""
    with np.errstate(over="ignore"):
        idx = np.arange(first, first + nsamples, dtype=U64)
        out = np.empty((nsamples, 2), dtype=np.float32)
        for c in (0, 1):
            s = np.full(nsamples, -262140, dtype=np.int64)
            for w in (0, 1):
                z = _mix64((U64(4) * idx + U64(2 * c + w + 1)) * U64(0x9E3779B97F4A7C15) + U64(seed))
                s += ((z & U64(0xFFFF)) + ((z >> U64(16)) & U64(0xFFFF))
                      + ((z >> U64(32)) & U64(0xFFFF)) + (z >> U64(48))).astype(np.int64)
            out[:, c] = s.astype(np.float32) * np.float32(2.0 ** -19)
    return out.reshape(-1)


def _gauss(seed: int, n: int, scale: float = 0.25) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * np.float32(scale)
    return x.view(np.float32)


def _ofdm32k(nsym: int) -> np.ndarray:
    """DVB-T2-like 32K OFDM: 27 265 active 256-QAM carriers, GI 1/128, x0.2 (dvbt2-blade.py:132)."""
    rng = np.random.default_rng(32768)
    nfft, nact, gi = 32768, 27265, 32768 // 128
    out = []
    lv = np.arange(-15, 16, 2, dtype=np.float64) / np.sqrt(170.0)
    for _ in range(nsym):
        spec = np.zeros(nfft, dtype=np.complex128)
        sym = lv[rng.integers(0, 16, nact)] + 1j * lv[rng.integers(0, 16, nact)]
        k = np.arange(nact) - nact // 2
        spec[k % nfft] = sym
        t = np.fft.ifft(spec) * np.sqrt(nfft) * np.sqrt(nfft / nact)
        out.append(np.concatenate([t[-gi:], t]))
    x = (np.concatenate(out) * 0.2).astype(np.complex64)
    return x.view(np.float32)


def _burst() -> np.ndarray:
    """Mostly weak noise with a few strong bursts: PAPR ~ 30 dB, hundreds of -g levels."""
    f = siggen(0, 200_000, 99).copy()
    f *= np.float32(2.0 ** -5)
    f[2 * 150_000: 2 * 150_064] *= np.float32(2.0 ** 6)
    return f


def _ties() -> np.ndarray:
    f = siggen(0, 4096, 5).copy()
    for k in (100, 400):       # duplicated power / component maxima: first occurrence must win
        f[2 * k] = 0.75
        f[2 * k + 1] = -0.75
    for k in (50, 3000):
        f[2 * k] = -0.75
        f[2 * k + 1] = 0.75
    return f


def _with(f: np.ndarray, k: int, re: float, im: float) -> np.ndarray:
    f = f.copy()
    f[2 * k] = re
    f[2 * k + 1] = im
    return f


def _negnan(f: np.ndarray, k: int) -> np.ndarray:
    f = f.copy()
    f.view(np.uint32)[2 * k + 1] = 0xFFC00000  # -NaN in Q
    return f


FIXTURES = {
    # name: callable returning the file image
    "appA_1M": lambda: siggen(0, 1 << 20, 1).tobytes(),           # SURVEY Appendix A known answer
    "appA_300k_s7": lambda: siggen(12345, 300_001, 7).tobytes(),  # ragged sample count, offset start
    "gauss_1M": lambda: _gauss(1234, 1 << 20).tobytes(),          # SURVEY §4.3 gauss_1M
    "const_4k": lambda: np.exp(1j * np.linspace(0, 50, 4096)).astype(np.complex64).tobytes(),
    "zeros_1000": lambda: np.zeros(2000, np.float32).tobytes(),
    "empty": lambda: b"",
    "one_sample": lambda: np.array([0.5, -0.25], np.float32).tobytes(),
    "odd_floats_short": lambda: siggen(0, 501, 3)[:1001].tobytes(),      # < 1 chunk: stale Q = 0.0
    "odd_floats_long": lambda: siggen(0, 9000, 3)[:17999].tobytes(),     # stale Q from previous chunk
    "odd_bytes_even": lambda: siggen(0, 9000, 4).tobytes()[: 8 * 8999 + 2],  # ragged, even floats
    "odd_bytes_odd": lambda: siggen(0, 9000, 4).tobytes()[: 8 * 8999 + 4 + 3],  # ragged + lone I
    "odd_bytes_short": lambda: siggen(0, 100, 4).tobytes()[: 8 * 77 + 4 + 1],
    "exact_chunks": lambda: siggen(0, 8192 * 3, 6).tobytes(),            # size % 64 KiB == 0
    "ties": lambda: _ties().tobytes(),
    "neg_only": lambda: (-np.abs(siggen(0, 5000, 8))).tobytes(),         # no positive component
    "nan": lambda: _with(siggen(0, 5000, 9), 1234, float("nan"), 0.1).tobytes(),
    "neg_nan_q": lambda: _negnan(siggen(0, 5000, 9), 777).tobytes(),
    "inf": lambda: _with(siggen(0, 5000, 9), 4321, float("inf"), 0.1).tobytes(),
    "ovf": lambda: _with(siggen(0, 5000, 9), 99, 3e38, 3e38).tobytes(),
    "denorm": lambda: (_gauss(77, 1 << 16) * np.float32(1e-22)).tobytes(),
    "ofdm32k": lambda: _ofdm32k(6).tobytes(),
    "burst": lambda: _burst().tobytes(),
}


def image(name: str) -> bytes:
    return FIXTURES[name]()


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()
