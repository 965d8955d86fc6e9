"""Throughput sweeps for DESIGN.md / profiles: channels 32..4096 at 2 Msps cu8 (config 3 curve),
the 8-channels-per-stream shape (config 2), and 10 Msps cs16 with 8 channels (config 5 shape)."""
import json
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from vdlm2dec_b200.api import Vdl2Gpu
from vdlm2dec_b200.synth_torch import make_device_workload

out = []
dev = torch.device("cuda")
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6650.0) if len(sys.argv) < 2 else float(sys.argv[1])


def run(nch, ns, cps=1, fmt="cu8", fs=2_000_000, sdrclk=500, reps=4, label=""):
    nstreams = nch // cps
    if fmt == "cu8":
        x, fos, nb = make_device_workload(nstreams, ns, seed=1000, device=dev)
    else:  # cs16 noise + a DC-free tone: timing only
        x = (torch.randn((nstreams, 2 * ns), device=dev) * 900).to(torch.int16)
        fos, nb = [(-50_000 - 125_000 * (c % 8)) for c in range(nstreams)], 0
    allfo = [f for f in range(-450_000, 475_000, 125_000) if abs(f) >= 50_000]
    if cps > 1:
        chans = [(c, 136_000_000 + allfo[c % cps], allfo[c % cps] if (c % cps) else fos[c // cps]) for c in range(nch)]
    else:
        chans = [(c, 136_975_000, fos[c]) for c in range(nch)]
    g = Vdl2Gpu(chans, fs=fs, sdrclk=sdrclk, fmt=fmt, ch_per_stream=cps, max_samples=ns, max_blocks=max(4096, 16 * nch))
    ms = []
    for _ in range(reps):
        g.process_device(x.data_ptr(), ns, x.stride(0))
        g.sync()
        ms.append(g.stats()["last_kernel_ms"])
        nblk = len(g.drain_blocks())
    t = float(np.median(ms[1:]))
    bps = x.element_size() * 2
    rec = {"label": label, "channels": nch, "ch_per_stream": cps, "format": fmt, "fs": fs, "samples_per_channel": ns, "kernel_ms": round(t, 4),
           "msamples_per_s": round(nch * ns / t / 1e3, 1), "hbm_gbs_algorithmic": round(nstreams * ns * bps / t / 1e6, 1),
           "frac_of_hbm_peak": round(nstreams * ns * bps / t / 1e6 / peak, 4), "blocks": nblk, "grid": g.stats()["grid"]}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    g.close()
    del x
    torch.cuda.empty_cache()


ns = 4_194_000
for nch in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    run(nch, ns if nch <= 2048 else ns // 2, label="config 3 curve: 1 channel per stream, 2 Msps cu8, bursts")
run(8, ns, cps=8, label="config 2: 8 channels from one 2 Msps cu8 stream")
run(1024, ns, cps=8, label="128 streams x 8 channels")
run(8, 10_000_000 // 1000 * 1600, cps=8, fmt="cs16", fs=10_000_000, sdrclk=2500, label="config 5 shape: 8 channels from one 10 Msps cs16 stream (noise)")
run(1024, 10_000_000 // 1000 * 400, cps=1, fmt="cs16", fs=10_000_000, sdrclk=2500, label="1024 x 10 Msps cs16, 1 channel per stream (noise)")
json.dump(out, open("gpurun_out/sweep_r1_v9.json", "w"), indent=1)
