"""Device-side IQ producers: synthetic captures in the reference's on-disk format
(headerless float32 I/Q pairs, gr_complex) generated on the GPU, so that >= 10 GiB inputs exist
without disk or PCIe traffic.

They restate the TAILS of the reference's transmit chains - what decides the envelope statistics `papr`
measures - with random constellation cells on the active carriers.  The upstream blocks (BCH / LDPC /
interleavers / pilots) only decide WHICH constellation points are sent and are out of scope.

  ofdm_symbols / ofdm_capture   dvbt2-blade.py:129-132,158-160: (tone reservation ->) IFFT -> cyclic prefix
                                (-> P1 preamble) -> x0.2 -> file_sink(gr.sizeof_gr_complex); 27 265 active
                                carriers of a 32K FFT, guard interval 1/128 (dvbt2rate.c:922-1010)
  dvbt_capture                  dvbt-blade.py:187-189,213-215: fft_vcc(carriers, inverse, unscaled) -> cyclic
                                prefix -> x0.0022097087; 1705 / 6817 carriers of a 2K / 8K FFT
  tr_tones / tr_kernel          the reserved carriers and the reference kernel of the tone-reservation step
                                (dvbt2_paprtr_cc, dvbt2-blade.py:52-54,129; EN 302 755 clause 9.6.2.1)

torch.fft (cuFFT) does the transforms: this is fixture plumbing, not the hot path; the tone-reservation
iterations run in the library's own kernel (papr_tr.cu).
"""
from __future__ import annotations

import math


def qam_levels(bits_per_axis: int = 4):
    """Amplitude levels of a square QAM axis normalised to unit average power per cell (256-QAM: 16 levels)."""
    m = 1 << bits_per_axis
    norm = math.sqrt(2.0 * (m * m - 1) / 3.0)
    return [(2.0 * i - (m - 1)) / norm for i in range(m)]


def tr_tones(fft_size: int = 32768, active: int = 27265, ntones: int = 288, seed: int = 5):
    """Reserved carriers (FFT bin numbers): `ntones` of the active carriers, spread pseudo-randomly like the
    standard's TR sets (32K: 288 cells per data symbol, dvbt2rate.c:1217); fixed by `seed`."""
    import numpy as np
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(active, ntones, replace=False))
    return ((pos - active // 2) % fft_size).astype(np.int32)


def tr_kernel(fft_size: int, tones, device="cuda"):
    """Reference kernel p: IFFT of 1 on every reserved tone, scaled so that p[0] = 1."""
    import torch
    spec = torch.zeros(fft_size, dtype=torch.complex64, device=device)
    spec[torch.as_tensor(tones, device=device).long()] = 1.0
    return (torch.fft.ifft(spec) * (fft_size / len(tones))).contiguous()


def ofdm_symbols(nsym: int, seed: int = 1, device="cuda", fft_size: int = 32768, active: int = 27265,
                 reserved=None, indices=None):
    """[nsym, fft_size] complex64 time-domain symbols with unit mean power: random 256-QAM cells on the active
    carriers (none on the `reserved` ones).  `indices` ([nsym, active, 2] integers 0..15) fixes the cells."""
    import torch
    if indices is None:
        g = torch.Generator(device=device).manual_seed(seed)
        indices = torch.randint(0, 16, (nsym, active, 2), generator=g, device=device)
    idx = torch.as_tensor(indices, device=device)
    lv = (2.0 * idx.to(torch.float32) - 15.0) / math.sqrt(170.0)        # 256-QAM, unit average power
    spec = torch.zeros((nsym, fft_size), dtype=torch.complex64, device=device)
    k = (torch.arange(active, device=device) - active // 2) % fft_size  # centred carriers
    spec[:, k] = torch.complex(lv[..., 0], lv[..., 1])
    used = active
    if reserved is not None:
        spec[:, torch.as_tensor(reserved, device=device).long()] = 0
        used = active - len(reserved)
    return (torch.fft.ifft(spec, dim=1) * (math.sqrt(fft_size) * math.sqrt(fft_size / used))).contiguous()


def p1_symbol(device="cuda", seed: int = 11):
    """A P1-like preamble (dvbt2-blade.py:131): a 1K OFDM symbol A with 384 of its carriers BPSK-modulated,
    framed by frequency-shifted copies of its two parts, C A B = 542 + 1024 + 482 samples, unit mean power."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    carriers = torch.sort(torch.randperm(853, generator=g, device=device)[:384]).values - 426
    bits = torch.randint(0, 2, (384,), generator=g, device=device).to(torch.float32) * 2.0 - 1.0
    spec = torch.zeros(1024, dtype=torch.complex64, device=device)
    spec[carriers % 1024] = torch.complex(bits, torch.zeros_like(bits))
    a = torch.fft.ifft(spec) * (1024.0 / math.sqrt(384.0))
    shift = torch.exp(2j * math.pi * torch.arange(1024, device=device) / 1024.0).to(torch.complex64)
    return torch.cat([a[:542] * shift[:542], a, a[542:] * shift[542:]])


def _tile(base, nsamples: int, device):
    """base [len] complex64 -> float32 tensor of 2*nsamples (I/Q interleaved), repeated with exact,
    power-preserving transforms (conjugate, swap I/Q, negate) up to `nsamples`."""
    import torch
    out = torch.empty(2 * nsamples, dtype=torch.float32, device=device)
    ov = out.view(-1, 2)
    b = torch.view_as_real(base.contiguous())                            # [len, 2] = I, Q
    variants = (b, b * torch.tensor([1.0, -1.0], device=device), b.flip(1), -b)
    pos, j = 0, 0
    while pos < nsamples:
        m = min(b.shape[0], nsamples - pos)
        ov[pos:pos + m] = variants[j % 4][:m]
        pos += m
        j += 1
    return out


def ofdm_capture(nsamples: int, seed: int = 1, device="cuda", fft_size: int = 32768, active: int = 27265,
                 guard: int = 128, scale: float = 0.2, base_symbols: int = 256, engine=None, tone_reservation: bool = False,
                 vclip: float = 3.3, iterations: int = 3, ntones: int = 288, p1: bool = False, indices=None):
    """float32 tensor of 2*nsamples (I/Q interleaved): a base block of `base_symbols` OFDM symbols generated
    with cuFFT - with `tone_reservation` corrected by the library's TR kernel (needs `engine`), with `p1` headed
    by a P1-like preamble - cyclic-prefixed, scaled by 0.2 and tiled up to `nsamples`."""
    import torch
    gi = fft_size // guard
    sym_len = fft_size + gi
    nsym = max(1, min(base_symbols, math.ceil(nsamples / sym_len)))
    tones = tr_tones(fft_size, active, ntones) if tone_reservation else None
    t = ofdm_symbols(nsym, seed, device, fft_size, active, reserved=tones, indices=indices)
    if tone_reservation:
        if engine is None:
            raise ValueError("tone_reservation needs the engine that owns the TR kernel")
        amax = 5.0 * ntones * math.sqrt(10.0 / (27.0 * active)) * math.sqrt(fft_size / active)
        engine.tr_reduce(t, tr_kernel(fft_size, tones, device), tones, vclip, iterations, amax)
    base = torch.cat([t[:, -gi:], t], dim=1).reshape(-1)                  # cyclic prefix, dvbt2-blade.py:130
    if p1:
        base = torch.cat([p1_symbol(device), base])                       # dvbt2-blade.py:131
    return _tile(base * scale, nsamples, device)                          # dvbt2-blade.py:132


def dvbt_capture(nsamples: int, seed: int = 1, device="cuda", mode: str = "8k", guard: int = 32, bits_per_axis: int = 3,
                 base_symbols: int = 256):
    """DVB-T (dvbt-blade.py:187-189): random 64-QAM cells on 1705 (2k) / 6817 (8k) carriers, the unscaled inverse
    FFT of GNU Radio's fft_vcc (N x ifft), cyclic prefix, x0.0022097087."""
    import torch
    fft_size, active = (2048, 1705) if mode == "2k" else (8192, 6817)
    gi = fft_size // guard
    nsym = max(1, min(base_symbols, math.ceil(nsamples / (fft_size + gi))))
    g = torch.Generator(device=device).manual_seed(seed)
    m = 1 << bits_per_axis
    lv = torch.tensor(qam_levels(bits_per_axis), dtype=torch.float32, device=device)
    idx = torch.randint(0, m, (nsym, active, 2), generator=g, device=device)
    spec = torch.zeros((nsym, fft_size), dtype=torch.complex64, device=device)
    k = (torch.arange(active, device=device) - active // 2) % fft_size
    spec[:, k] = torch.complex(lv[idx[..., 0]], lv[idx[..., 1]])
    t = torch.fft.ifft(spec, dim=1) * float(fft_size)                     # fft_vcc(..., forward=False): no 1/N
    base = torch.cat([t[:, -gi:], t], dim=1).reshape(-1)
    return _tile(base * 0.0022097087, nsamples, device)                   # dvbt-blade.py:189


def write_cfile(path: str, iq) -> None:
    """blocks.file_sink(gr.sizeof_gr_complex, path) equivalent (dvbt2-blade.py:158-160)."""
    iq.detach().cpu().numpy().tofile(path)
