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


def check(x, chans, fs, sdrclk, fmt, cps, ns, nchk=2):
    """parity of the swept point itself: `nchk` channels of the timed tensor through a fresh handle vs the CPU checker"""
    from oracle import pyoracle
    from tests.parity_util import oracle_kind
    sub = x[:max(1, nchk // cps)].contiguous()
    sel = list(range(sub.shape[0] * cps))
    g = Vdl2Gpu([chans[c] for c in sel], fs=fs, sdrclk=sdrclk, fmt=fmt, ch_per_stream=cps, max_samples=ns)
    g.process_device(sub.data_ptr(), ns, sub.stride(0) * sub.element_size())
    g.sync()
    blocks = g.drain_blocks()
    host = sub.cpu().numpy()
    nb = 0
    for c in sel:
        chn, Fr, Fo = chans[c]
        want = pyoracle.Oracle(oracle_kind(), chn=chn, Fr=Fr, Fo=Fo, fs=fs, sdrclk=sdrclk, taps=pyoracle.TAP_BLOCKS).feed(host[c // cps], fmt).blocks
        want = want[want["end_dump"] < ns // (fs // 1000) * 84]
        got = blocks[blocks["chn"] == chn]
        assert len(want) == len(got) and np.array_equal(want["data"], got["data"]) and np.array_equal(want["sync_dump"], got["sync_dump"]), (fs, c)
        nb += len(want)
    return nb


def run(nch, ns, cps=1, fmt="cu8", fs=2_000_000, sdrclk=500, reps=4, label="", noise=False, parity=0):
    nstreams = nch // cps
    if fs == 2_000_000:
        raster = [f for f in range(-450_000, 475_000, 125_000) if abs(f) >= 50_000]
    else:  # eight carriers spread over the band, on the 25 kHz raster the oscillator table needs, clear of DC and of the band edges
        span = (fs // 2 - 100_000) // 25_000 * 25_000
        raster = [f // 25_000 * 25_000 for f in np.linspace(-span, span, 8).astype(int)]
    if noise:  # cs16 noise: timing only
        x = (torch.randn((nstreams, 2 * ns), device=dev) * 900).to(torch.int16)
        fos, nb = [raster[c % len(raster)] for c in range(nch)], 0
    else:
        short = ns < fs
        x, fos, nb = make_device_workload(nstreams, ns, seed=1000, device=dev, fs=fs, fmt=fmt, fos=raster, ch_per_stream=cps,
                                          amp=(12.0, 18.0) if (cps > 1 and fmt == "cu8") else (25.0, 70.0),
                                          first_burst=0.01 if short else 0.25, gap=(0.02, 0.05) if short else (0.15, 0.6))
    chans = [(c, 136_975_000 if cps == 1 else 136_000_000 + (fos[c] % 1_000_000), fos[c]) for c in range(nch)]
    g = Vdl2Gpu(chans, fs=fs, sdrclk=sdrclk, fmt=fmt, ch_per_stream=cps, max_samples=ns, max_blocks=max(4096, 16 * nch))
    ms = []
    for _ in range(reps):
        g.process_device(x.data_ptr(), ns, x.stride(0) * x.element_size())
        g.sync()
        ms.append(g.stats()["last_kernel_ms"])
        nblk = len(g.drain_blocks())
    t = float(np.median(ms[1:]))
    bps = x.element_size() * 2
    pb = check(x, chans, fs, sdrclk, fmt, cps, ns, parity) if parity else None
    rec = {"label": label, "parity_blocks_checked": pb, "window_samples": round(fs / 84000, 1), "channels": nch, "ch_per_stream": cps, "format": fmt, "fs": fs, "samples_per_channel": ns, "kernel_ms": round(t, 4),
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
def run_f3(nstreams, cps, ns, reps=4):
    """row f3: the one-pass channeliser (decimated streams of all channels of every stream to HBM) next to the fused kernel's phase-1 share"""
    raster = [f for f in range(-450_000, 475_000, 125_000) if abs(f) >= 50_000]
    x, fos, nb = make_device_workload(nstreams, ns, seed=1000, device=dev, fos=raster, ch_per_stream=cps, amp=(12.0, 18.0))
    nch = nstreams * cps
    chans = [(c, 136_000_000 + fos[c] % 1_000_000, fos[c]) for c in range(nch)]
    g = Vdl2Gpu(chans, ch_per_stream=cps, max_samples=ns)
    rows = ns // 2000
    outb = torch.empty((nch, rows * 84), dtype=torch.complex64, device=dev)
    ms = []
    for _ in range(reps):
        g.channelise_device(x.data_ptr(), ns, x.stride(0), outb.data_ptr(), outb.stride(0))
        g.sync()
        ms.append(g.stats()["last_kernel_ms"])
    t = float(np.median(ms[1:]))
    bytes_in, bytes_out = nstreams * ns * 2, nch * rows * 84 * 8
    rec = {"label": f"row f3: one-pass channeliser, {nstreams} streams x {cps} channels, 2 Msps cu8 -> 84 ksps complex float", "channels": nch, "ch_per_stream": cps,
           "samples_per_channel": ns, "kernel_ms": round(t, 4), "msamples_per_s": round(nch * ns / t / 1e3, 1), "bytes_in": bytes_in, "bytes_out": bytes_out,
           "hbm_gbs_algorithmic": round((bytes_in + bytes_out) / t / 1e6, 1), "frac_of_hbm_peak": round((bytes_in + bytes_out) / t / 1e6 / peak, 4)}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    g.close()
    del x, outb
    torch.cuda.empty_cache()


run_f3(128, 8, ns)
run_f3(1024, 1, ns)
run_f3(1, 8, ns)
# BASELINE config 5 ("Airspy 10 Msps cs16, 8 channels, FIR-tap length sweep 64 -> 512"): the reference's channel filter is the boxcar over
# one dump, fs / 84000 samples long, so the sweep with an oracle is the rate sweep fs = 84 kHz x L
for L in (75, 125, 250, 500):
    fs = 84_000 * L
    rows = 1600 if L <= 125 else (800 if L == 250 else 400)
    run(8, fs // 1000 * rows, cps=8, fmt="cs16", fs=fs, sdrclk=fs // 4000, parity=8, label=f"config 5: 8 channels from one {fs / 1e6:g} Msps cs16 stream, window {L}, bursts")
    run(256, fs // 1000 * (400 if L <= 125 else 100), cps=1, fmt="cs16", fs=fs, sdrclk=fs // 4000, parity=2,
        label=f"config 5 rate at HBM scale: 256 x {fs / 1e6:g} Msps cs16, 1 channel per stream, window {L}, bursts")
run(1024, 10_000_000 // 1000 * 400, cps=1, fmt="cs16", fs=10_000_000, sdrclk=2500, noise=True, label="1024 x 10 Msps cs16, 1 channel per stream (noise)")
json.dump(out, open("gpurun_out/sweep_r2.json", "w"), indent=1)
