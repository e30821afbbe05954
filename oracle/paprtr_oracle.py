"""TEST INFRASTRUCTURE — numpy restatement of the tone-reservation PAPR reduction the reference configures
(drmpeg/dtv-utils dvbt2-blade.py:52-54,129: dtv.dvbt2_paprtr_cc, vclip 3.3, 3 iterations).

PARITY UNPINNED: the block is GNU Radio gr-dtv's (dvbt2_paprtr_cc_impl.cc), which /root/reference only calls and
does not contain, and the reference holds no vectors for it.  What is restated is the published algorithm,
EN 302 755 V1.4.1 clause 9.6.2.1 ("PAPR reduction algorithm"), in float64, symbol by symbol, with plain
Python / numpy loops.  Only tests/ may import this file; nothing under dtv-utils_b200/ does."""
import numpy as np


def tr_kernel(fft_size, tones):
    spec = np.zeros(fft_size, np.complex128)
    spec[np.asarray(tones)] = 1.0
    return np.fft.ifft(spec) * (fft_size / len(tones))  # p[0] = 1


def tr_reduce(x, p, tones, vclip, iterations, amax):
    """x: [nsym, N] complex (time-domain symbols).  Returns (corrected symbols, reserved-tone values, iterations)."""
    x = np.array(x, np.complex128)
    nsym, n = x.shape
    tones = np.asarray(tones, np.int64)
    r = np.zeros((nsym, len(tones)), np.complex128)
    iters = np.zeros(nsym, np.int32)
    for s in range(nsym):
        it = 0
        while it < iterations:
            mag2 = x[s].real ** 2 + x[s].imag ** 2
            m = int(np.argmax(mag2))  # first maximum
            y = np.sqrt(mag2[m])
            if not y > vclip * (1.0 + 4e-6):  # the kernel works in float32: a peak it has just clipped to Vclip
                break                          # reads back as Vclip to ~1e-7 relative and must not count as a new one
            u = x[s, m] / y
            v = u * np.exp(-2j * np.pi * ((tones * m) % n) / n)
            rv = r[s] * np.conj(v)
            alpha_k = np.sqrt(np.maximum(amax * amax - rv.imag ** 2, 0.0)) + rv.real
            alpha = min(y - vclip, float(alpha_k.min()))
            if not alpha > 0:
                break
            x[s] -= alpha * u * p[(np.arange(n) - m) % n]
            r[s] -= alpha * v
            it += 1
        iters[s] = it
    return x, r, iters
