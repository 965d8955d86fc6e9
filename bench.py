#!/usr/bin/env python
"""bench.py -- IQ Msamples/s of multi-channel VDL2 D8PSK demodulation on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is ONE pass of the fused front-end kernel over one batch: `--channels` (1024) independent
2 Msps cu8 IQ streams x `--samples` (2^22) samples per GPU, synthetic (AWGN + seeded valid VDL2
bursts), resident in HBM before the timed region.  Channels shard across GPUs with no data-path
collective (weak scaling: 1024 channels per GPU); torch.distributed is used only for the
barrier and the max-over-ranks of the device time.

One JSON line on stdout (rank 0).  `value` = channel-samples/s with inputs resident in HBM,
`e2e` = the same through the C ABI with pinned HOST buffers (H2D + block drain inside the timed
region), `roofline` = algorithmic HBM bytes / CUDA-event kernel time against the measured copy
bandwidth, `cpu_baseline` = the reference's own d8psk.c path (oracle/_ref) on the host cores.
`--impl reference` times only that CPU path, on all host threads, same config/metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s (multi-ch D8PSK demod)"
FS = 2_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=1024, help="channels per GPU (BASELINE config 3: 1024)")
    ap.add_argument("--samples", type=int, default=1 << 22, help="IQ samples per channel per step")
    ap.add_argument("--ch-per-stream", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 5],
                    help="BASELINE.json config: 3 (default; config 4 = the same per GPU at N=8) 1024 x 2 Msps cu8, one channel per stream; "
                         "2 = 8 channels from ONE 2 Msps cu8 stream (the rtl.c shape); 5 = 8 channels from one cs16 stream at the "
                         "Airspy-class rate 10.5 Msps, plus the window-length sweep fs = 84 kHz x {75, 125, 250, 500}")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle self-check of the timed workload")
    ap.add_argument("--check-channels", type=int, default=8)
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.2)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        if not rows:  # nvidia-smi was too slow to start: one direct query right after the timed region
            try:
                q = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
                rows = [[x.strip() for x in q[0].split(",")]] if q else []
            except Exception:
                rows = []
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU legs
def _pick_cpu_kind():
    """oracle/_ref (the reference's own d8psk.c, -Ofast) when it loads and runs here, else the port."""
    from oracle import pyoracle
    for kind in ("ref_native", "ref_fast"):
        if not pyoracle.available(kind):
            continue
        probe = ("import sys; sys.path.insert(0, %r); import numpy as np; from oracle import pyoracle; "
                 "pyoracle.time_cu8(%r, np.full(40000, 127, np.uint8), 1)") % (ROOT, kind)
        if subprocess.run([sys.executable, "-c", probe], capture_output=True).returncode == 0:  # SIGILL-safe
            return kind, "reference"
    if not pyoracle.available("port"):
        pyoracle.build("port")
    return "port", "port"


def cpu_throughput(iq_rows, fos, target_cpu_seconds: float, threads: int | None = None):
    """Reference CPU path on `threads` host threads, one private channel each (throughput mode of
    BASELINE.md section 3): returns (Msamples/s, threads, kind, sample description)."""
    import numpy as np
    from oracle import pyoracle
    kind, label = _pick_cpu_kind()
    threads = threads or os.cpu_count() or 1
    n = iq_rows.shape[1] // 2
    # calibrate one pass on one thread, then size reps so total CPU work ~ target_seconds
    t1 = pyoracle.time_cu8(kind, iq_rows[0], 1, Fo=fos[0])
    reps = max(1, min(256, int(round(target_cpu_seconds / threads / max(t1, 1e-6)))))
    res = [0.0] * threads

    def work(i):
        r = i % iq_rows.shape[0]
        res[i] = pyoracle.time_cu8(kind, iq_rows[r], reps, Fo=fos[r])

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    total = threads * reps * n
    lib = {"ref_native": "reference d8psk.c -Ofast -march=native", "ref_fast": "reference d8psk.c -Ofast -march=x86-64-v3",
           "port": "oracle port -O2"}[kind]
    desc = f"{threads} threads x {reps} passes x {n} samples of one channel each ({lib}); single-thread {n / t1 / 1e6:.1f} Msamples/s"
    return total / wall / 1e6, threads, label, desc


def host_workload(nrows: int, nsamples: int, seed: int):
    """The first `nrows` channels of the GPU arm's workload, built on the HOST by the same generator and seed
    (vdlm2dec_b200.synth_torch.make_device_workload on the CPU device: same burst library, positions, amplitudes, Fo
    per channel; the noise comes from the CPU generator instead of the CUDA one)."""
    import torch
    from vdlm2dec_b200.synth_torch import make_device_workload
    x, fos, _nb = make_device_workload(nrows, nsamples, seed=seed, device=torch.device("cpu"), group=4)
    return x.numpy(), fos[:nrows]


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ns = args.samples // 2000 * 2000
    # one private channel per host thread, full step length (no cache-resident replay): rows 0..threads-1 of the GPU workload
    iq, fos = host_workload(min(threads, 64), ns, seed=1000)
    vals = []
    kind = desc = None
    walls = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, threads, kind, desc = cpu_throughput(iq, fos, target_cpu_seconds=1.0 * threads)  # ~1 s wall per step
        if s >= args.warmup:
            vals.append(v)
            walls.append(time.perf_counter() - t0)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),  # one step = one bounded sample (config.sample), not the GPU arm's step
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.channels} channels/GPU x {ns} samples, 2 Msps cu8 IQ, 1 ch/stream (BASELINE config 3; "
                               f"config 4 = the same per GPU at N=8)", "channels_per_gpu": args.channels, "samples_per_channel": ns,
                   "sample": desc, "inputs": "channels 0.. of the GPU arm's workload (same generator, seed 1000), one per host thread, "
                                             "each the full step length"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_near_gpu(local_rank):
    """N > 1 only: confine this rank's host threads to the CPUs NVML reports as nearest to its GPU while the pinned e2e
    buffer is allocated and fed, so that its pages sit on the GPU's NUMA node (8 ranks x 50 GB/s of H2D is more than one
    socket's memory should serve across the inter-socket link).  Returns the previous affinity (to restore) or None."""
    try:
        import pynvml
        import torch
        old = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except TypeError:  # older bindings want bytes
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        if not os.sched_getaffinity(0):
            os.sched_setaffinity(0, old)
            return None
        return old
    except Exception:
        return None


def self_check(x, chans, fos, ns, cps, device, nchk, fs=FS, sdrclk=500, fmt="cu8"):
    """`nchk` channels of the timed tensor (spread over the channel range) through a fresh handle, against the CPU checker
    (oracle/_ref = the reference's d8psk.c when present, else the port) fed the same bytes: completed blocks bit exact."""
    import numpy as np
    from oracle import pyoracle
    from vdlm2dec_b200.api import Vdl2Gpu
    kind = "port"
    if pyoracle.available("ref"):
        try:
            pyoracle.load("ref")
            kind = "ref"
        except OSError:
            pass
    if kind == "port" and not pyoracle.available("port"):
        pyoracle.build("port")
    nstreams = x.shape[0]
    pick = sorted({int(i * (nstreams - 1) / max(1, nchk - 1)) for i in range(nchk)})
    sub = x[pick].contiguous()
    sel = [c for s_ in pick for c in range(s_ * cps, (s_ + 1) * cps)]
    g = Vdl2Gpu([chans[c] for c in sel], fs=fs, sdrclk=sdrclk, fmt=fmt, ch_per_stream=cps, device=device, max_samples=ns)
    g.process_device(sub.data_ptr(), ns, sub.stride(0) * sub.element_size())
    g.sync()
    blocks = g.drain_blocks()
    host = sub.cpu().numpy()
    nblk, bad = 0, []
    for i, c in enumerate(sel):
        chn, Fr, Fo = chans[c]
        o = pyoracle.Oracle(kind, chn=chn, Fr=Fr, Fo=Fo, fs=fs, sdrclk=sdrclk, taps=pyoracle.TAP_BLOCKS).feed(host[i // cps], fmt)
        want = o.blocks
        want = want[want["end_dump"] < ns // (fs // 1000) * 84]
        got = blocks[blocks["chn"] == chn]
        same = len(want) == len(got) and all(
            a["sync_dump"] == b["sync_dump"] and a["end_dump"] == b["end_dump"] and a["nbrow"] == b["nbrow"] and a["nlbyte"] == b["nlbyte"]
            and np.array_equal(a["data"], b["data"]) for a, b in zip(want, got))
        nblk += len(want)
        if not same:
            bad.append(int(chn))
    return {"ok": not bad and nblk > 0, "channels": len(sel), "blocks": int(nblk), "mismatching_channels": bad,
            "checker": "reference d8psk.c (oracle/_ref, -O2)" if kind == "ref" else "oracle port",
            "what": "completed blocks (trigger/end position, nbrow, nlbyte, data[8][255]) bit exact"}


def config5_sweep(dev, local_rank, peak, steps):
    """BASELINE config 5's "FIR-tap length sweep 64 -> 512".  The reference has no taps: its channel filter is the boxcar over one
    dump, fs / 84000 samples long (d8psk.c:374-381), so the sweep that has an oracle is the rate sweep fs = 84 kHz x L with fs a
    multiple of 25 kHz: L = 75, 125, 250, 500 (6.3 / 10.5 / 21 / 42 Msps), 8 channels from one cs16 stream, every point checked
    against the CPU checker on the timed tensor."""
    import numpy as np
    import torch
    from vdlm2dec_b200.api import Vdl2Gpu
    from vdlm2dec_b200.synth_torch import make_device_workload
    out = []
    for L in (75, 125, 250, 500):
        fs = 84_000 * L
        ns = fs // 1000 * (1600 if L <= 125 else (800 if L == 250 else 400))
        span = (fs // 2 - 100_000) // 25_000 * 25_000
        raster = [int(f) // 25_000 * 25_000 for f in np.linspace(-span, span, 8)]
        x, fos, nb = make_device_workload(1, ns, seed=500 + L, device=dev, fs=fs, fmt="cs16", fos=raster, ch_per_stream=8,
                                          first_burst=0.01, gap=(0.02, 0.05))
        chans = [(c, 136_000_000 + fos[c] % 1_000_000, fos[c]) for c in range(8)]
        g = Vdl2Gpu(chans, fs=fs, sdrclk=fs // 4000, fmt="cs16", ch_per_stream=8, device=local_rank, max_samples=ns, max_blocks=(steps + 4) * max(nb, 64) + 4096)
        ms = []
        for _ in range(steps + 1):
            g.process_device(x.data_ptr(), ns, x.stride(0) * x.element_size())
            g.sync()
            ms.append(g.stats()["last_kernel_ms"])
            g.drain_blocks()
        t = float(np.median(ms[1:]))
        par = self_check(x, chans, fos, ns, 8, local_rank, 1, fs=fs, sdrclk=fs // 4000, fmt="cs16")
        out.append({"window_samples": L, "fs": fs, "samples": ns, "bursts": nb, "kernel_ms": t, "msamples_per_s": 8 * ns / t / 1e3,
                    "stream_gbs": ns * 4 / t / 1e6, "frac_of_hbm_peak": ns * 4 / t / 1e6 / peak, "parity_ok": par["ok"], "parity_blocks": par["blocks"]})
        g.close()
        del x
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from vdlm2dec_b200.api import OPT_OVERLAP, Vdl2Gpu
    from vdlm2dec_b200.synth_torch import make_device_workload

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    fs, sdrclk, fmt, bps = FS, 500, "cu8", 2
    if args.config == 2:      # the reference's own use: up to 8 frequencies from one 2 MHz stream (README.md:4, vdlm2.h:26)
        args.channels, args.ch_per_stream = 8, 8
    elif args.config == 5:    # Airspy-class rate, cs16 (extension; the reference's Airspy mode is float32 real at 5/6 Msps, air.c:123,134-138)
        args.channels, args.ch_per_stream = 8, 8
        fs, sdrclk, fmt, bps = 10_500_000, 2625, "cs16", 4
        args.samples = min(args.samples * 4, 1 << 24)
    nch, ns = args.channels, args.samples // (fs // 1000) * (fs // 1000)
    cps = args.ch_per_stream
    nstreams = nch // cps
    t_gen = time.time()
    if fs == FS:
        raster = [f for f in range(-450_000, 475_000, 125_000) if abs(f) >= 50_000]
    else:
        span = (fs // 2 - 100_000) // 25_000 * 25_000
        raster = [int(f) // 25_000 * 25_000 for f in np.linspace(-span, span, 8)]
    x, fos, nbursts = make_device_workload(nstreams, ns, seed=1000 + 17 * rank, device=dev, fs=fs, fmt=fmt, fos=raster, ch_per_stream=cps,
                                           amp=(12.0, 18.0) if (cps > 1 and fmt == "cu8") else (25.0, 70.0))
    chans = [(c + rank * nch, 136_975_000 if cps == 1 else 136_000_000 + fos[c] % 1_000_000, fos[c]) for c in range(nch)]
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    # OPT_OVERLAP: back-to-back launches may overlap on the device (programmatic dependent launch); the library then
    # records no per-launch events, the bench brackets with its own
    g = Vdl2Gpu(chans, fs=fs, sdrclk=sdrclk, fmt=fmt, ch_per_stream=cps, device=local_rank, max_samples=ns, taps=OPT_OVERLAP,
                max_blocks=(args.steps + 4) * max(nbursts, 64) + 4096)
    stream = torch.cuda.ExternalStream(g.cuda_stream, device=dev)

    def step():
        g.process_device(x.data_ptr(), ns, x.stride(0) * x.element_size())

    # ---- warm-up (also validates: every placed burst must come back as a block -- checked below on the FIRST step,
    #      which starts from a fresh state; later steps re-feed the buffer with carried state, so their seams add events)
    blocks_seen, blocks_first = 0, None
    for _ in range(max(1, args.warmup)):
        step()
        g.sync()
        blocks_seen = len(g.drain_blocks())
        if blocks_first is None:
            blocks_first = blocks_seen
    # ---- timed region: K launches, CUDA events on the launching stream, barrier + sync both sides
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(40):  # wait until nvidia-smi delivers (its start-up can take a second on a fresh box)
        if sampler.rows:
            break
        time.sleep(0.05)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = g.stats()["kernel_launches"]
    w0 = time.time()
    e0.record(stream)
    kernel_ms = []
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    w1 = time.time()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(w0, w1)
    launches = g.stats()["kernel_launches"] - launches0
    blocks_timed = len(g.drain_blocks())
    # per-launch kernel time (events recorded by the library around the kernel on its stream)
    for _ in range(3):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        step()
        k1.record(stream)
        g.sync()
        kernel_ms.append(k0.elapsed_time(k1))
        g.drain_blocks()
    kms = sum(kernel_ms) / len(kernel_ms)

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    units = nch * ns * args.steps * world  # channel-samples over all ranks
    value = units / (ms_max * 1e-3) / 1e6

    # ---- self-check of what was timed: channels of the timed tensor through a fresh handle vs the CPU oracle, blocks bit exact
    parity = None
    if rank == 0 and not args.no_check:
        parity = self_check(x, chans, fos, ns, cps, local_rank, args.check_channels, fs=fs, sdrclk=sdrclk, fmt=fmt)
        if not parity["ok"]:
            raise SystemExit("bench.py: the timed workload does NOT decode like the oracle: " + json.dumps(parity))
    if blocks_first < nbursts - max(2, nbursts // 200):   # a burst cut by the end of the buffer may be missing
        raise SystemExit(f"bench.py: {nbursts} bursts placed per step but only {blocks_first} blocks came back from the first step")

    # ---- end to end through the C ABI with pinned host buffers (H2D + drain inside the timed region)
    e2e = None
    if not args.no_e2e:
        old_affinity = bind_near_gpu(local_rank) if world > 1 else None
        hx = torch.empty((nstreams, 2 * ns), dtype=x.dtype, pin_memory=True)
        hx.copy_(x)
        torch.cuda.synchronize()
        nrep = max(2, min(args.steps, 4))
        g.process_ptr(hx.data_ptr(), ns, hx.stride(0) * hx.element_size())
        g.drain_blocks()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(nrep):
            g.process_ptr(hx.data_ptr(), ns, hx.stride(0) * hx.element_size())  # returns after the kernel finished
            d2h += g.drain_blocks().nbytes + 32
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": nch * ns * nrep * world / float(tt.item()) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(hx.numel() * hx.element_size()), "d2h_bytes_per_step": int(d2h // nrep), "steps": nrep,
               "host_numa_bound": old_affinity is not None}
        del hx
        if old_affinity is not None:
            os.sched_setaffinity(0, old_affinity)

    # ---- block pipeline behind the demodulator: the blocks of one step through
    #      vdl2_link_kernel (RS + HDLC + FCS), next to the reference's blk_thread path on one host core
    link = None
    if rank == 0:
        try:
            step()
            g.sync()
            blks = g.drain_blocks()
            lms, nfr = [], 0
            for _ in range(4):
                fr, _st, _ = g.link_decode(blks, want_rows=False)
                lms.append(g.stats()["last_link_ms"])
                nfr = len(fr)
            lk = sum(lms[1:]) / len(lms[1:])
            step()
            g.sync()
            t0 = time.perf_counter()
            fr2, _b2 = g.drain_frames()
            fixed_s = time.perf_counter() - t0
            # rows f1 + f4 end to end: ordered, packed frames + field records straight from the device (page-locked buffers)
            g.drain_frames_packed()   # allocates the page-locked buffers of the mirror once
            pk = []
            for _ in range(3):
                step()
                g.sync()
                t0 = time.perf_counter()
                hdrs, pdata, recs = g.drain_frames_packed()
                pk.append((time.perf_counter() - t0, g.last_pack_ms, g.stats()["last_link_ms"], len(hdrs), len(pdata)))
            pk_host, pk_ms, pk_link_ms, pk_n, pk_bytes = min(pk)
            lbytes = len(blks) * 2080 + nfr * 2048
            link = {"kernel": "vdl2_link_kernel", "blocks": int(len(blks)), "frames": int(nfr), "kernel_ms": lk,
                    "blocks_per_s": len(blks) / (lk * 1e-3) if lk > 0 else None,
                    "algorithmic_bytes": lbytes, "achieved_gbs": lbytes / (lk * 1e-3) / 1e9 if lk > 0 else None,
                    "bound": "lsu (shared-memory table look-ups; 4 KB of HBM traffic per block)",
                    "drain_frames_ms_host": pk_host * 1e3, "drain_frames_what": "vdl2_drain_frames_packed: block pipeline + ranking + packing "
                    "+ field records on the device, three page-locked copies, host time of the whole call",
                    "drain_frames_fixed_records_ms_host": fixed_s * 1e3, "frames_fused": int(pk_n), "packed_bytes": int(pk_bytes)}
            abytes = pk_n * (2048 + 48 + 32) + pk_bytes
            avlc = {"kernel": "vdl2_frame_rank + _scan + _pack + vdl2_avlc_kernel (one CUDA-event bracket)", "frames": int(pk_n),
                    "kernel_ms": pk_ms, "frames_per_s": pk_n / (pk_ms * 1e-3) if pk_ms > 0 else None, "algorithmic_bytes": int(abytes),
                    "achieved_gbs": abytes / (pk_ms * 1e-3) / 1e9 if pk_ms > 0 else None,
                    "bound": "latency (a few thousand frames per step: four launches of tens of microseconds each)",
                    "acars_frames": int((recs["kind"] == 2).sum()) if recs is not None else None}
            if not args.no_cpu:
                from oracle import pyoracle
                akind = "ref" if pyoracle.out_available("ref") else "port"
                # random payloads that look like XID groups (0x82) send the reference's outxid() into an endless loop on a negative
                # 16-bit group length (outxid.c:268-299): keep them out of the CPU leg
                keep = np.flatnonzero(recs["kind"] != 1)[:2048]
                sub_n = len(keep)
                pdata_pad = np.concatenate([pdata, np.zeros(64, np.uint8)])
                t1 = pyoracle.out_time(akind, pdata_pad, hdrs["offset"][keep], hdrs["len"][keep], 1)
                reps = max(1, min(50, int(2.0 / max(t1, 1e-4))))
                tt = pyoracle.out_time(akind, pdata_pad, hdrs["offset"][keep], hdrs["len"][keep], reps)
                avlc["cpu_baseline"] = {"value": sub_n * reps / tt, "unit": "frames/s", "cores": 1, "kind": "reference" if akind == "ref" else "port",
                                        "sample": f"{reps} passes over {sub_n} frames of the step through the reference's out() in JSON mode "
                                                  f"(out.c + outacars.c + outxid.c + label.c + cJSON.c; it also FORMATS, which the device record "
                                                  f"leaves to the host)" if akind == "ref" else f"{reps} passes over {sub_n} frames through the port's field walk"}
            link["avlc"] = avlc
            if not args.no_cpu:
                from oracle import pyoracle
                kind = "ref" if pyoracle.link_available("ref") else "port"
                sub = blks[:min(len(blks), 2048)]
                t1 = pyoracle.link_time(kind, sub, 1)
                reps = max(1, min(50, int(3.0 / max(t1, 1e-4))))
                tt = pyoracle.link_time(kind, sub, reps)
                link["cpu_baseline"] = {"value": len(sub) * reps / tt, "unit": "blocks/s", "cores": 1,
                                        "kind": "reference" if kind == "ref" else "port",
                                        "sample": f"{reps} passes over {len(sub)} blocks of the step through the reference's blk_thread "
                                                  f"(vdlm2.c + rs.c + crc.c, single consumer thread as in the reference)"}
        except Exception as exc:  # the headline must not depend on the optional row
            link = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the (only) kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = nstreams * ns * bps  # 2 B per cu8 (4 B per cs16) IQ sample per STREAM; output bytes are negligible
    kms_region = ms_total / args.steps  # rank 0, CUDA events on the launching stream over the timed region
    achieved = alg_bytes / (kms_region * 1e-3) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("channels") == nch and tr.get("ch_per_stream") == cps:
            traffic = tr["dram_bytes_per_sample_stream"] * nstreams * ns
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s",
                "kernel": "vdl2_frontend_kernel", "kernel_ms": kms_region, "kernel_ms_isolated": kms, "algorithmic_bytes_per_launch": alg_bytes}

    cpu = None
    if not args.no_cpu and world == 1 and fmt == "cu8":
        nrows = min(4, nstreams)
        sub = x[:nrows, : 2 * min(ns, 1 << 21)].cpu().numpy()
        v, cores, kind, desc = cpu_throughput(sub, fos[:nrows], args.cpu_seconds)
        cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": desc}

    sweep = config5_sweep(dev, local_rank, peak, 3) if args.config == 5 else None
    line = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{nch} channels/GPU x {ns} samples, {fs / 1e6:g} Msps {fmt} IQ, {cps} ch/stream (BASELINE config "
                               + {3: "3; config 4 = the same per GPU at N=8)", 2: "2: 8 channels from one stream, the rtl.c shape)",
                                  5: "5 at window length 125; see `sweep`)"}[args.config],
                   "baseline_config": args.config, "channels_per_gpu": nch, "samples_per_channel": ns,
                   "bytes_per_step_per_gpu": alg_bytes,
                   "l2": "input per step (8 GiB at defaults) >> 126 MB L2; no flush needed" if alg_bytes > (1 << 29) else
                         "input per step fits the 126 MB L2: the shared stream is read once from HBM and cps times from L2 by design "
                         "(FP32/tensor issue bound, not HBM bound; the HBM fraction is reported for completeness)",
                   "bursts_per_step_per_gpu": nbursts, "blocks_decoded_per_step": blocks_timed // max(1, args.steps) if blocks_timed else blocks_seen,
                   "parallelism": f"channels sharded, {world} GPU(s), no collective", "gen_seconds": round(t_gen, 1),
                   "launch_mode": "back-to-back launches with programmatic dependent launch (tails overlap)"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "parity_checked": parity, "blocks_first_step": blocks_first, "link": link, "sweep": sweep,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
