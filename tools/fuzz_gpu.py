"""Randomized sweep of the CUDA path THROUGH THE C ABI on a B200 against the CPU checker (the reference's own d8psk.c from
oracle/_ref when present, else the port): random Fo on the 25 kHz raster, amplitude, noise level, burst spacing and payload
length per case, for each of the three 8-bit mixers (int8 tensor cores = default, IDP.4A, generic fp32).  Cases are batched as
the channels of one handle (64 per launch).  Per case tests/parity_util.compare_channel checks blocks, trigger positions and
timing, symbol positions, hard decisions and the 1e-5 rad soft-symbol bar; the tool also reports the worst |dD|, the number of
Gray-index moves and every case outside the bar.

    python tools/fuzz_gpu.py <seed> <cases> [lowsnr] [--json out.json]
`lowsnr`: amplitude 3-12 LSB in noise of sigma 6-24 LSB (marginal and false triggers all over).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from tests.parity_util import compare_channel, oracle_kind, run_oracle, wrap_diff
from vdlm2dec_b200 import synth
from vdlm2dec_b200.api import OPT_DP4A_MIX, OPT_FLOAT_MIX, TAP_DUMPS, TAP_SYMS, TAP_SYNCS, Vdl2Gpu

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lowsnr = "lowsnr" in sys.argv[3:]
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
rng = np.random.default_rng(seed0)
fos = [f for f in range(-450_000, 475_000, 25_000) if abs(f) >= 50_000]
n = 800_000
BATCH = 64
MIXERS = (("mma", 0), ("dp4a", OPT_DP4A_MIX), ("float", OPT_FLOAT_MIX))
summary = {m: dict(cases=0, outside_bar=0, decision_mismatch=0, max_dD=0.0, max_dump_relerr=0.0, gi_moves=0, symbols=0, blocks=0, syncs=0)
           for m, _ in MIXERS}
bad_cases = []
t0 = time.time()
done = 0
while done < N:
    nb = min(BATCH, N - done)
    cases, iqs = [], []
    for _ in range(nb):
        fo = int(rng.choice(fos))
        seed = int(rng.integers(0, 1 << 30))
        if lowsnr:
            amp_lo = float(rng.uniform(3, 12))
            sigma = float(rng.choice([6.0, 8.0, 12.0, 16.0, 24.0]))
        else:
            amp_lo = float(rng.uniform(8, 60))
            sigma = float(rng.choice([0.0, 2.0, 4.0, 8.0]))
        period = int(rng.integers(25_000, 80_000))
        spec = synth.standard_channel(seed=seed, nsamples=n, Fo=fo, period=period, payload_bytes=(14, 600), amp=(amp_lo, amp_lo * 1.5),
                                      noise_sigma=sigma)
        cases.append(dict(fo=fo, seed=seed, amp=amp_lo, sigma=sigma, period=period))
        iqs.append(synth.render_channel(spec, n))
    iq = np.stack(iqs)
    oracles = [run_oracle(iq[c], cases[c]["fo"], chn=c) for c in range(nb)]
    for mname, opt in MIXERS:
        g = Vdl2Gpu([(c, 136_975_000, cases[c]["fo"]) for c in range(nb)], taps=TAP_DUMPS | TAP_SYNCS | TAP_SYMS | opt, max_samples=n)
        g.process(iq)
        blocks = g.drain_blocks()
        S = summary[mname]
        for c in range(nb):
            o = oracles[c]
            gd, gs, gy = g.read_dumps(c), g.read_syncs(c), g.read_syms(c)
            S["cases"] += 1
            S["blocks"] += len(o.blocks)
            S["syncs"] += len(o.syncs)
            S["symbols"] += len(gy)
            # the raw numbers first (they do not depend on the bar) ...
            osy = o.syms
            if len(osy) == len(gy) and len(gy):
                S["max_dD"] = max(S["max_dD"], float(wrap_diff(osy["D"], gy["D"]).max()))
                S["gi_moves"] += int((osy["gi"] != gy["gi"]).sum())
            od = o.dumps[:len(gd)]
            S["max_dump_relerr"] = max(S["max_dump_relerr"], float(np.abs(od - gd).max() / (np.sqrt(np.mean(np.abs(od) ** 2)) + 1e-30)))
            # ... then the verdict of the parity bar
            try:
                compare_channel(o, blocks[blocks["chn"] == c], gs, gy, gd, None, ndump_limit=len(gd))
            except AssertionError as e:
                msg = str(e)[:200]
                only_phase = msg.startswith("soft symbol deviates")
                S["outside_bar"] += 1
                if not only_phase:
                    S["decision_mismatch"] += 1
                bad_cases.append(dict(mixer=mname, **cases[c], error=msg))
                print("OUTSIDE", mname, cases[c], msg, flush=True)
        del g
    done += nb
    print(f"{done}/{N} cases, {time.time() - t0:.0f} s", {m: (s["outside_bar"], f"{s['max_dD']:.2e}") for m, s in summary.items()}, flush=True)
res = dict(seed=seed0, cases=N, mode="lowsnr" if lowsnr else "normal", checker=oracle_kind(), samples_per_case=n, mixers=summary, outside=bad_cases,
           seconds=round(time.time() - t0, 1))
print(json.dumps(res))
if out_json:
    json.dump(res, open(out_json, "w"), indent=1)
