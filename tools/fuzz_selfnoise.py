"""How far apart are the REFERENCE's own two builds on the fuzz cases of tools/fuzz_gpu.py?  oracle/_ref/libvdl2ref_O2.so (strict
IEEE, the parity oracle) vs libvdl2ref_fast.so (-Ofast, the project's own flags, CMakeLists.txt:4), same seeds, same generator.
Reports per-symbol |dD| beyond 1e-5 rad: the floor any implementation in different fp32 rounding sits on.  CPU only.
    python tools/fuzz_selfnoise.py <seed> <cases> [lowsnr]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.pyoracle import Oracle
from tests.parity_util import wrap_diff
from vdlm2dec_b200 import synth
seed0 = int(sys.argv[1]); N = int(sys.argv[2]); lowsnr = "lowsnr" in sys.argv[3:]
rng = np.random.default_rng(seed0)
fos = [f for f in range(-450_000, 475_000, 25_000) if abs(f) >= 50_000]
n = 800_000
outside = decisions = 0
worst = 0.0
nsym = 0
for it in range(N):
    fo = int(rng.choice(fos)); seed = int(rng.integers(0, 1 << 30))
    if lowsnr:
        amp_lo = float(rng.uniform(3, 12)); sigma = float(rng.choice([6.0, 8.0, 12.0, 16.0, 24.0]))
    else:
        amp_lo = float(rng.uniform(8, 60)); sigma = float(rng.choice([0.0, 2.0, 4.0, 8.0]))
    period = int(rng.integers(25_000, 80_000))
    spec = synth.standard_channel(seed=seed, nsamples=n, Fo=fo, period=period, payload_bytes=(14, 600), amp=(amp_lo, amp_lo * 1.5), noise_sigma=sigma)
    iq = synth.render_channel(spec, n)
    a = Oracle("ref", Fo=fo).feed(iq); b = Oracle("ref_fast", Fo=fo).feed(iq)
    sa, sb = a.syms, b.syms
    ba, bb = a.blocks, b.blocks
    if (len(sa) != len(sb) or not np.array_equal(sa["dump"], sb["dump"]) or len(ba) != len(bb) or not np.array_equal(ba["data"], bb["data"])
            or not np.array_equal(ba["sync_dump"], bb["sync_dump"]) or not np.array_equal(sa["v"] > 0.5, sb["v"] > 0.5)):
        decisions += 1
        continue
    nsym += len(sa)
    if len(sa):
        d = float(wrap_diff(sa["D"], sb["D"]).max())
        worst = max(worst, d)
        outside += d >= 1e-5
print(dict(seed=seed0, cases=N, mode="lowsnr" if lowsnr else "normal", cases_with_a_symbol_beyond_1e5_rad=int(outside), cases_with_different_events=decisions,
           max_dD=worst, symbols=nsym))
