"""Wall time of the reference program on one capture (BASELINE configs 1 and 2: 1 and 8 channels from one 2 Msps cu8
stream, replayed through the UNMODIFIED rtl.c callback by the file-backed fake librtlsdr): all-CPU binary, the binary with
our d8psk.o, and the one whose d8psk.o also replaces vdlm2.o + rs.o (row f1).  Not a bench line: the drop-in protocol
(one 32768-sample block per barrier round, 8 B/sample Cbuff) is what it is; this shows what the swap buys as is."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from vdlm2dec_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINS = {k: os.path.join(ROOT, "oracle", "_ref", v) for k, v in
        (("reference", "vdlm2dec_cpu"), ("gpu d8psk.o", "vdlm2dec_gpu"), ("gpu d8psk.o + block pipeline", "vdlm2dec_gpu_link"))}


def capture(path, fos, nblk):
    n = 32768 * nblk
    x = np.zeros(n, dtype=np.complex128)
    for i, fo in enumerate(fos):
        spec = synth.standard_channel(seed=50 + i, nsamples=n - 60_000, Fo=fo, period=200_000, payload_bytes=(20, 300),
                                      amp=(20.0, 28.0) if len(fos) > 1 else (40.0, 60.0), noise_sigma=0.0)
        x += synth.render_channel(spec, n, fmt="cf32").astype(np.float64).view(np.complex128)
    rng = np.random.default_rng(1)
    x += 4.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    synth.quantise(x, "cu8").tofile(path)
    return n


if __name__ == "__main__":
    import re
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 1024     # 1024 blocks = 2^25 samples = 16.8 s of signal
    REP = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    out = []
    freqs8 = ["136.975", "136.850", "136.725", "136.800", "136.650", "136.775", "136.900", "136.675"]
    with tempfile.TemporaryDirectory() as d:
        fmax = max(float(f) for f in freqs8)
        fos = [int(round((float(f) - fmax) * 1e6)) - 50_000 for f in freqs8]
        cap1 = os.path.join(d, "cap.cu8")
        capture(cap1, fos, nblk)
        cap = os.path.join(d, "cap_full.cu8")       # the same capture REP times back to back
        data = open(cap1, "rb").read()
        with open(cap, "wb") as g:
            for _ in range(REP):
                g.write(data)
        n = 32768 * nblk * REP
        for freqs in (freqs8[:1], freqs8):
            for name, b in BINS.items():
                if not os.path.exists(b):
                    continue
                best = None
                for _ in range(2):
                    t0 = time.perf_counter()
                    p = subprocess.run([b, "-G", "-E", "-U", "-r", "0", *freqs], env=dict(os.environ, VDL2_FAKE_IQ=cap, VDL2_SHIM_STATS="1"),
                                       capture_output=True)
                    wall = time.perf_counter() - t0
                    # our objects time themselves from the first feed to the last block delivered (CUDA context creation, about
                    # 1.5 s, is start-up, not throughput); the all-reference program has no start-up to speak of: wall clock
                    m = re.search(rb"fed (\d+) samples in ([0-9.]+) s", p.stderr)
                    dt = float(m.group(2)) if m else wall
                    if best is None or dt < best[0]:
                        best = (dt, wall, int(m.group(1)) if m else n, bool(m))
                lines = p.stdout.count(b"[#")
                dt, wall, fed, own = best
                out.append({"channels": len(freqs), "binary": name, "samples": fed, "seconds": round(dt, 3), "wall_seconds": round(wall, 3),
                            "timer": "program's own (first feed to last block delivered)" if own else "wall clock of the whole program",
                            "messages": lines, "stream_msps": round(1e-6 * fed / dt, 1), "channel_msps": round(len(freqs) * 1e-6 * fed / dt, 1)})
                print(json.dumps(out[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dropin_bench.json"), "w"), indent=1)
