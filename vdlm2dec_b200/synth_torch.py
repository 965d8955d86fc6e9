"""Device-side synthesis of the bench workload (torch, not part of the hot path).

Builds `nch` independent 2 Msps cu8 IQ streams directly in HBM: AWGN plus seeded valid VDL2
bursts (the same transmit chain as synth.py, whose bit/symbol stages run on the host because
they are tiny; only pulse shaping, mixing and quantisation run on the GPU).  Returns the
number of bursts placed so the bench can check that every one of them was decoded.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import synth


def _rc(x: torch.Tensor, alpha: float = 0.6) -> torch.Tensor:
    den = 1.0 - (2.0 * alpha * x) ** 2
    sing = den.abs() < 1e-9
    den = torch.where(sing, torch.ones_like(den), den)
    y = torch.sinc(x) * torch.cos(math.pi * alpha * x) / den
    lim = (math.pi / 4.0) * float(np.sinc(1.0 / (2.0 * alpha)))
    return torch.where(sing, torch.full_like(y, lim), y)


def burst_waveform(phase_idx: np.ndarray, fs: float, frac: float, device) -> torch.Tensor:
    """complex64 baseband burst; symbol k peaks at sample (6 + k) * T + frac."""
    T = fs / synth.SYMRATE
    K = len(phase_idx)
    L = int(math.ceil((K + 12) * T))
    n = torch.arange(L, dtype=torch.float64, device=device)
    u = (n - frac) / T - 6.0
    kc = torch.floor(u).to(torch.int64)
    ph = torch.from_numpy(np.exp(1j * (np.pi / 4.0) * phase_idx.astype(np.float64))).to(device)
    out = torch.zeros(L, dtype=torch.complex128, device=device)
    for j in range(-4, 6):
        k = kc + j
        ok = (k >= 0) & (k < K)
        kk = k.clamp(0, K - 1)
        out += torch.where(ok, ph[kk] * _rc(u - k.to(torch.float64)), torch.zeros_like(out))
    return out.to(torch.complex64)


class BurstLibrary:
    def __init__(self, seed: int, device, fs: int = 2_000_000, n: int = 24, payload_bytes=(30, 600)):
        rng = np.random.default_rng(seed)
        self.fs = fs
        self.items = []
        for _ in range(n):
            nb = int(rng.integers(payload_bytes[0], payload_bytes[1] + 1))
            b = synth.make_burst(rng, nb)
            pidx = synth.burst_phase_indices(b, rng=rng)
            w = burst_waveform(pidx, fs, float(rng.uniform(0, fs / synth.SYMRATE)), device)
            self.items.append((b, w))


def make_device_workload(nch: int, nsamples: int, seed: int, device, fs: int = 2_000_000, fos=None,
                         noise_sigma: float = 8.0, gap=(0.15, 0.6), group: int = 16, lib: BurstLibrary | None = None,
                         fmt: str = "cu8", first_burst=None, ch_per_stream: int = 1, amp=(25.0, 70.0)):
    """-> (uint8 tensor [nch, 2*nsamples] for cu8, int16 for cs16 (the same signal x 64), list of Fo per channel, number of bursts placed).
    first_burst: latest start of a channel's first burst in seconds (default 0.25).
    ch_per_stream > 1: `nch` counts STREAMS; every stream carries the bursts of ch_per_stream channels at fos[0 .. ch_per_stream)
    (the rtl "8 frequencies from one 2 MHz stream" shape, README.md:4); the returned Fo list is per stream-major channel."""
    fos = fos or [f for f in range(-450_000, 475_000, 125_000) if abs(f) >= 50_000]
    lib = lib or BurstLibrary(seed, device, fs)
    out = torch.empty((nch, 2 * nsamples), dtype=torch.uint8 if fmt == "cu8" else torch.int16, device=device)
    rng = np.random.default_rng(seed + 1)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 2)
    nb_total = 0
    cps = ch_per_stream
    ch_fo = [fos[c % len(fos)] for c in range(nch)] if cps == 1 else [fos[k] for _ in range(nch) for k in range(cps)]
    for c0 in range(0, nch, group):
        g = min(group, nch - c0)
        x = torch.randn((g, nsamples), dtype=torch.complex64, device=device, generator=gen) * (noise_sigma * math.sqrt(2.0))
        for i, k in ((i, k) for i in range(g) for k in range(cps)):
            fo = ch_fo[(c0 + i) * cps + k]
            t = int(rng.uniform(fs // 1000, (first_burst or 0.25) * fs))
            while True:
                b, w = lib.items[int(rng.integers(0, len(lib.items)))]
                L = w.numel()
                if t + L + 2 * (fs // 1000) >= nsamples:
                    break
                a_ = float(rng.uniform(*amp))
                cfo = float(rng.uniform(-500.0, 500.0))
                n = torch.arange(t, t + L, dtype=torch.float64, device=device)
                cyc = (n * ((fo + cfo) / fs)) % 1.0
                ang = (cyc * (2.0 * math.pi) + float(rng.uniform(0, 2 * math.pi))).to(torch.float32)
                rot = torch.complex(torch.cos(ang), torch.sin(ang))
                x[i, t:t + L] += a_ * w * rot
                nb_total += 1
                t += L + int(rng.uniform(*gap) * fs)
        iq = torch.view_as_real(x)  # [g, ns, 2]
        if fmt == "cu8":
            q = (iq + 127.37).round_().clamp_(0, 255).to(torch.uint8)
        else:
            q = (iq * 64.0).round_().clamp_(-32768, 32767).to(torch.int16)
        out[c0:c0 + g] = q.reshape(g, 2 * nsamples)
        del x, iq, q
    return out, ch_fo, nb_total
