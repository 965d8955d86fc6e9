"""Randomized pin of the oracle PORT against the REFERENCE's own d8psk.c + viterbi.c compiled in place (oracle/_ref): random Fo on
the 25 kHz raster, input format, amplitude, noise, burst spacing, chunking of the feed; every tap (T1 dumps .. T6 blocks) compared
bit for bit.  Needs /root/reference-built oracle/_ref; CPU only.
    python tools/fuzz_oracle.py <seed> <cases> [rates]
`rates`: Airspy real-sample mode at 5 and 6 Msps and complex cs16 at 10 Msps instead of 2 Msps (Fo over the wider raster).
Round 1: 2 400 cases at 2 Msps (seeds 21-24 x 600) and 1 000 with `rates` (seeds 41-44 x 250): every tap bit identical."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle.pyoracle import Oracle
from tests.test_oracle import _taps_equal
from vdlm2dec_b200 import synth

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
RATES = len(sys.argv) > 3 and sys.argv[3] == "rates"
fos = [f for f in range(-450_000, 475_000, 25_000) if abs(f) >= 50_000]
bad = 0
t0 = time.time()


def other_rate_case(it):
    """One case at 5 / 6 Msps float32 real (air.c) or 10 Msps cs16; returns True when every tap is bit identical."""
    fs, fmt, real = [(6_000_000, "f32real", True), (5_000_000, "f32real", True), (10_000_000, "cs16", False)][int(rng.integers(0, 3))]
    lim = fs // 4 - 100_000 if real else fs // 2 - 100_000
    fo = int(rng.choice([f for f in range(-lim, lim + 1, 25_000) if abs(f) >= 50_000]))
    n, seed = 1_500_000, int(rng.integers(0, 1 << 30))
    spec = synth.standard_channel(seed=seed, nsamples=n - 100_000, Fo=fo, fs=fs, period=int(0.02 * fs), payload_bytes=(14, 300),
                                  amp=(20.0, 50.0), noise_sigma=float(rng.choice([0.0, 2.0, 6.0])))
    x = synth.render_channel(spec, n, fs=fs, fmt=fmt)
    if real:
        x = (x.astype(np.float32) / 64).astype(np.float32)
    kw = dict(Fo=fo, fs=fs, sdrclk=fs // 4000, real_input=real)
    r, p = Oracle("ref", **kw).feed(x, fmt), Oracle("port", **kw).feed(x, fmt)
    try:
        _taps_equal(r, p)
        print(it, "ok", fs, fmt, fo, "blocks", len(r.blocks), flush=True)
        return True
    except AssertionError as e:
        print(it, "MISMATCH", dict(fs=fs, fmt=fmt, fo=fo, seed=seed), str(e)[:200], flush=True)
        return False


for it in range(N):
    if RATES:
        bad += not other_rate_case(it)
        continue
    fo = int(rng.choice(fos))
    seed = int(rng.integers(0, 1 << 30))
    fmt = str(rng.choice(["cu8", "cu8", "cs8", "cs16", "cf32"]))
    n = 600_000
    amp = float(rng.uniform(6, 60))
    sigma = float(rng.choice([0.0, 2.0, 4.0, 8.0, 16.0]))
    spec = synth.standard_channel(seed=seed, nsamples=n, Fo=fo, period=int(rng.integers(25_000, 80_000)), payload_bytes=(14, 600),
                                  amp=(amp, amp * 1.5), noise_sigma=sigma)
    iq = synth.render_channel(spec, n, fmt=fmt)
    r, p = Oracle("ref", Fo=fo), Oracle("port", Fo=fo)
    per = iq.size if rng.random() < 0.5 else int(rng.integers(1000, 200_000)) * 2
    for k in range(0, iq.size, per):
        r.feed(iq[k:k + per], fmt)
        p.feed(iq[k:k + per], fmt)
    try:
        _taps_equal(r, p)
        print(it, "ok", fo, fmt, "sigma", sigma, "blocks", len(r.blocks), "syncs", len(r.syncs), flush=True)
    except AssertionError as e:
        bad += 1
        print(it, "MISMATCH", dict(fo=fo, seed=seed, fmt=fmt, amp=amp, sigma=sigma, per=per), str(e)[:200], flush=True)
print("done", N, "bad", bad, "in", round(time.time() - t0, 1), "s")
