"""Randomized pins behind the demodulator, CPU only:
  * block pipeline: oracle port vs the reference's vdlm2.c + rs.c + crc.c compiled in place, on mixed blocks (clean, correctable,
    beyond the code, garbage, multi-frame, stuffing-heavy) -- frames, rs() results, frame counts and corrected rows;
  * frame fields (row f4): the kernel's per-frame walk compiled for the host (tests/emul) vs the oracle port, records byte for byte.
    python tools/fuzz_link_avlc.py [first_seed] [seeds]
Round 1: seeds 200-229 -> 12 000 blocks / 8 605 frames identical; seeds 100-119 -> 20 000 frames identical."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import pyoracle
from tests import emul
from tests.link_util import make_blocks
from tests.test_avlc_oracle import _frame_records
from tests.test_link_oracle import _same

first = int(sys.argv[1]) if len(sys.argv) > 1 else 200
count = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kinds = ("clean", "errors", "heavy", "garbage", "multi", "stuff")
t0, bad, nb, nf = time.time(), 0, 0, 0
for seed in range(first, first + count):
    blocks = make_blocks(seed, 400, kinds)
    fr, sr, rr = pyoracle.link_decode("ref", blocks)
    fp, sp, rp = pyoracle.link_decode("port", blocks)
    nb += len(blocks)
    nf += len(fp)
    try:
        _same(fp, sp, rp, fr, sr, rr, nbytes=False)
    except AssertionError as e:
        bad += 1
        print("LINK MISMATCH seed", seed, str(e)[:100])
print(f"link port vs reference: {nb} blocks, {nf} frames, {bad} bad seeds, {time.time() - t0:.1f} s")
t0, bad, nfr = time.time(), 0, 0
for seed in range(first, first + count):
    rec, frames = _frame_records(n_acars=200, n_other=800, seed=seed)
    got = emul.avlc(rec)
    want = np.array([pyoracle.avlc_extract(f) for f in frames], dtype=pyoracle.AVLC_DT)
    nfr += len(frames)
    if got.tobytes() != want.tobytes():
        bad += 1
        print("AVLC MISMATCH seed", seed)
print(f"frame-field walk vs port: {nfr} frames, {bad} bad seeds, {time.time() - t0:.1f} s")
