"""Row f3 probe: the one-pass channeliser alone (phase 1 without the demodulator) -- timing and an ncu target.
    python tools/f3_probe.py [streams] [channels per stream] [samples]"""
import sys
import torch
sys.path.insert(0, ".")
from vdlm2dec_b200.api import Vdl2Gpu
from vdlm2dec_b200.synth_torch import make_device_workload

nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ns = (int(sys.argv[3]) if len(sys.argv) > 3 else 4_194_000) // 2000 * 2000
dev = torch.device("cuda")
raster = [f for f in range(-450_000, 475_000, 125_000) if abs(f) >= 50_000]
x, fos, nb = make_device_workload(nstreams, ns, seed=1000, device=dev, fos=raster, ch_per_stream=cps, amp=(12.0, 18.0) if cps > 1 else (25.0, 70.0))
nch = nstreams * cps
g = Vdl2Gpu([(c, 136_975_000, fos[c]) for c in range(nch)], ch_per_stream=cps, max_samples=ns)
rows = ns // 2000
out = torch.empty((nch, rows * 84), dtype=torch.complex64, device=dev)
for r in range(4):
    g.channelise_device(x.data_ptr(), ns, x.stride(0), out.data_ptr(), out.stride(0))
    g.sync()
    ms = g.stats()["last_kernel_ms"]
    print(f"rep {r}: {ms:.3f} ms  in {nstreams * ns * 2 / ms / 1e6:.0f} GB/s  in+out {(nstreams * ns * 2 + nch * rows * 84 * 8) / ms / 1e6:.0f} GB/s")
