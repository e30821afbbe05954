"""Device-side IQ producers: synthetic captures in the reference's on-disk format
(headerless float32 I/Q pairs, gr_complex) generated on the GPU, so that >= 10 GiB inputs exist
without disk or PCIe traffic.

`ofdm_capture` restates the tail of the reference's DVB-T2 transmit chain, dvbt2-blade.py:129-132,158-160
(IFFT -> cyclic prefix -> x0.2 -> file_sink(gr.sizeof_gr_complex)), with random 256-QAM cells on the
27 265 active carriers of a 32K FFT and guard interval 1/128 (dvbt2rate.c:922-1010 for the carrier
count).  The upstream blocks (BCH/LDPC/interleavers/pilots/P1) only decide WHICH constellation
points are sent, not the envelope statistics `papr` measures, and are out of scope.  torch.fft
(cuFFT) does the transform: this is fixture plumbing, not the hot path.
"""
from __future__ import annotations

import math


def ofdm_capture(nsamples: int, seed: int = 1, device="cuda", fft_size: int = 32768, active: int = 27265,
                 guard: int = 128, scale: float = 0.2, base_symbols: int = 256):
    """float32 tensor of 2*nsamples (I/Q interleaved).  A base block of `base_symbols` OFDM symbols
    is generated with cuFFT and tiled with exact, power-preserving transforms (conjugate, swap I/Q,
    negate) up to `nsamples`."""
    import torch
    gi = fft_size // guard
    sym_len = fft_size + gi
    nsym = max(1, min(base_symbols, math.ceil(nsamples / sym_len)))
    g = torch.Generator(device=device).manual_seed(seed)
    idx = torch.randint(0, 16, (nsym, active, 2), generator=g, device=device)
    lv = (2.0 * idx.to(torch.float32) - 15.0) / math.sqrt(170.0)        # 256-QAM, unit average power
    spec = torch.zeros((nsym, fft_size), dtype=torch.complex64, device=device)
    k = (torch.arange(active, device=device) - active // 2) % fft_size  # centred carriers
    spec[:, k] = torch.complex(lv[..., 0], lv[..., 1])
    t = torch.fft.ifft(spec, dim=1) * (math.sqrt(fft_size) * math.sqrt(fft_size / active) * scale)
    base = torch.cat([t[:, -gi:], t], dim=1).reshape(-1).contiguous()    # cyclic prefix, dvbt2-blade.py:130
    out = torch.empty(2 * nsamples, dtype=torch.float32, device=device)
    ov = out.view(-1, 2)
    b = torch.view_as_real(base)                                         # [len, 2] = I, Q
    variants = (b, b * torch.tensor([1.0, -1.0], device=device), b.flip(1), -b)
    pos, j = 0, 0
    while pos < nsamples:
        m = min(b.shape[0], nsamples - pos)
        ov[pos:pos + m] = variants[j % 4][:m]
        pos += m
        j += 1
    return out


def write_cfile(path: str, iq) -> None:
    """blocks.file_sink(gr.sizeof_gr_complex, path) equivalent (dvbt2-blade.py:158-160)."""
    iq.detach().cpu().numpy().tofile(path)
