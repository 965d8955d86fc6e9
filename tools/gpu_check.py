"""First-light check on a B200: CUDA path vs the CPU oracle, verbose.  Run under gpurun."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
from tests.parity_util import make_channels, run_oracle, compare_channel
from vdlm2dec_b200.api import Vdl2Gpu, TAP_DUMPS, TAP_STEPS, TAP_SYNCS, TAP_SYMS

def main():
    nch, ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    fmt = sys.argv[3] if len(sys.argv) > 3 else "cu8"
    specs, iq = make_channels(nch, ns, seed=5, fmt=fmt)
    chans = [(c, 136_975_000, specs[c].Fo) for c in range(nch)]
    g = Vdl2Gpu(chans, fmt=fmt, taps=TAP_DUMPS | TAP_STEPS | TAP_SYNCS | TAP_SYMS, max_samples=ns)
    t = time.time(); g.process(iq); print("gpu process", time.time() - t, g.stats())
    blocks = g.drain_blocks()
    print("gpu blocks", len(blocks))
    bad = 0
    for c in range(nch):
        o = run_oracle(iq[c], specs[c].Fo, fmt=fmt, chn=c)
        gd, gs, gy, gt = g.read_dumps(c), g.read_syncs(c), g.read_syms(c), g.read_steps(c)
        lim = len(gd)
        try:
            rep = compare_channel(o, blocks[blocks["chn"] == c], gs, gy, gd, gt, ndump_limit=lim)
            print("ch", c, "OK", rep, "bursts", len(specs[c].bursts))
        except AssertionError as e:
            bad += 1
            print("ch", c, "FAIL", e)
            print("   oracle: dumps", len(o.dumps), "steps", len(o.steps), "syncs", o.syncs[:4], "nsyms", len(o.syms), "blocks", len(o.blocks))
            print("   gpu   : dumps", len(gd), "steps", len(gt), "syncs", gs[:4], "nsyms", len(gy), "blocks", (blocks["chn"] == c).sum())
            if len(gd):
                od = o.dumps[:len(gd)]
                dd = np.abs(od - gd)
                print("   dumps maxdiff", dd.max(), "at", dd.argmax(), "rms", np.sqrt(np.mean(np.abs(od)**2)), od[:3], gd[:3])
                print("   row-wise maxdiff", [float(dd[i*84:(i+1)*84].max()) for i in range(min(6, len(gd)//84))])
            if len(gt) and len(o.steps):
                n = min(len(gt), len(o.steps))
                print("   steps dump eq", np.array_equal(o.steps["dump"][:n], gt["dump"][:n]), "P maxdiff", np.abs(o.steps["P"][:n]-gt["P"][:n]).max())
    print("RESULT", "PASS" if bad == 0 else f"FAIL({bad})")
    return bad

if __name__ == "__main__":
    try:
        sys.exit(1 if main() else 0)
    except Exception:
        traceback.print_exc(); sys.exit(2)
