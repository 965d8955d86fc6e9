"""Generates tests/golden/link_golden.npz from the REFERENCE link layer (vdlm2.c + rs.c + crc.c compiled in
place -> oracle/_ref/libvdl2linkref.so).  Run in the build container, where /root/reference is mounted:
    python -m tests.golden.make_link_golden"""
import os

import numpy as np

from oracle import pyoracle
from tests.link_util import make_blocks

if __name__ == "__main__":
    pyoracle.build("all")
    blocks = make_blocks(77, 48)
    f, s, rows = pyoracle.link_decode("ref", blocks)
    out = os.path.join(os.path.dirname(__file__), "link_golden.npz")
    np.savez_compressed(out, blocks=np.frombuffer(blocks.tobytes(), np.uint8), frame_len=f["len"], frame_block=f["block"],
                        frame_bytes=np.concatenate([x["hdata"][:x["len"]] for x in f]), rs=s["rs"], rows_after=rows)
    print(out, len(blocks), "blocks", len(f), "frames", os.path.getsize(out), "bytes")
