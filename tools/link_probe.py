"""Device-time probe of the block pipeline kernel (vdl2_link_kernel) on synthetic blocks."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests.link_util import make_blocks
from vdlm2dec_b200.api import Vdl2Gpu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kinds = ("clean", "errors") if len(sys.argv) > 2 and sys.argv[2] == "clean" else ("clean", "errors", "heavy", "garbage", "multi", "stuff")
base = make_blocks(11, 512, kinds)
blocks = np.tile(base, (n + 511) // 512)[:n]
g = Vdl2Gpu([(0, 136_975_000, -50_000)], max_samples=200_000)
for r in range(4):
    f, s, _ = g.link_decode(blocks, want_rows=False)
    ms = g.stats()["last_link_ms"]
    print(f"rep {r}: {ms:.4f} ms  {n / ms / 1e3:.2f} M blocks/s  frames {len(f)}  rows repaired {(s['rs'] > 0).sum()}  given up {(s['rs'] < 0).sum()}")
