"""Prints the measured parity margins of the CUDA path against the CPU oracle (B200): integer mixer vs fp32 mixer."""
import sys
sys.path.insert(0, ".")
from tests.parity_util import compare_channel, make_channels, run_oracle
from vdlm2dec_b200.api import OPT_FLOAT_MIX, TAP_DUMPS, TAP_STEPS, TAP_SYMS, TAP_SYNCS, Vdl2Gpu

nch, n = 8, 1_600_000
for fmt in ("cu8", "cs8"):
    specs, iq = make_channels(nch, n, seed=21, fmt=fmt)
    chans = [(c, 136_975_000, specs[c].Fo) for c in range(nch)]
    for name, opt in (("integer mixer", 0), ("fp32 mixer", OPT_FLOAT_MIX)):
        g = Vdl2Gpu(chans, fmt=fmt, taps=TAP_DUMPS | TAP_STEPS | TAP_SYNCS | TAP_SYMS | opt, max_samples=n)
        g.process(iq)
        blocks = g.drain_blocks()
        worst = {}
        for c in range(nch):
            o = run_oracle(iq[c], specs[c].Fo, fmt=fmt, chn=c)
            gd = g.read_dumps(c)
            rep = compare_channel(o, blocks[blocks["chn"] == c], g.read_syncs(c), g.read_syms(c), gd, g.read_steps(c), ndump_limit=len(gd))
            for k, v in rep.items():
                if isinstance(v, float):
                    worst[k] = max(worst.get(k, 0.0), v)
                elif k == "gi_flips":
                    worst[k] = worst.get(k, 0) + v
        print(f"{fmt} {name:14s} blocks {len(blocks):3d}  " + "  ".join(f"{k}={v:.3g}" for k, v in sorted(worst.items())))
        g.close()
