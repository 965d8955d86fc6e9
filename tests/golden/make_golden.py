"""Generate tests/golden/*.npz from the REFERENCE build (oracle/_ref, i.e. the reference's own
d8psk.c compiled in place).  Run in the build container where /root/reference is mounted:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.pyoracle import Oracle, table  # noqa: E402
from vdlm2dec_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    np.savez_compressed(os.path.join(HERE, "tables.npz"), sync=table("ref", 0), mflt=table("ref", 1),
                        soft1=table("ref", 2), soft2=table("ref", 3), soft3=table("ref", 4))
    Fo, n = -125_000, 400_000
    spec = synth.standard_channel(seed=77, nsamples=n, Fo=Fo, period=30_000, payload_bytes=(20, 120))
    iq = synth.render_channel(spec, n)
    o = Oracle("ref", Fo=Fo).feed(iq)
    assert len(o.blocks) == len(spec.bursts) >= 2
    np.savez_compressed(os.path.join(HERE, "burst_ref.npz"), iq=iq, Fo=Fo, dumps_sub=o.dumps[::97], syncs=o.syncs,
                        sym_D=o.syms["D"], sym_gi=o.syms["gi"], blocks=o.blocks)
    print("golden written:", len(o.blocks), "blocks,", len(o.syms), "symbols")


if __name__ == "__main__":
    main()
