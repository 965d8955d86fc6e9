"""First-run check of the row-f4 kernel on a GPU without importing torch or pytest's GPU probe (seconds, not a minute):
records through the C ABI against the oracle port, byte for byte.  Prints AVLC_GPU_OK or the first mismatch."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import pyoracle
from tests.test_avlc_oracle import _frame_records
from vdlm2dec_b200 import api

t0 = time.time()
rec, frames = _frame_records(n_acars=100, n_other=400)
g = api.Vdl2Gpu([(0, 136_975_000, -50_000)], max_samples=200_000)
got = g.avlc_extract(rec)
g.close()
want = np.array([pyoracle.avlc_extract(f) for f in frames], dtype=pyoracle.AVLC_DT)
if got.tobytes() == want.tobytes():
    print(f"AVLC_GPU_OK {len(rec)} frames, {time.time() - t0:.1f} s")
else:
    bad = [i for i in range(len(rec)) if got[i].tobytes() != want[i].tobytes()]
    print("AVLC_GPU_MISMATCH", len(bad), "first", bad[0], got[bad[0]], want[bad[0]])
    sys.exit(1)
