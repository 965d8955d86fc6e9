"""Program-level timing of row f2 (file replay, vdlm2dec_b200/csrc/file_shim.c) on one 2 Msps cu8 capture, 1 and 8 channels.
The replay binaries (the reference's unmodified main.c + our objects) report their own steady state with -v: the time from
the first batch read to the last batch delivered, handle creation excluded ("Replayed N samples in T s").  The two
binaries that keep the rtl.c callback protocol (all-reference, and reference + our d8psk.o) are timed by wall clock on the
same capture through the fake dongle.  Not a bench line; writes gpurun_out/replay_bench.json."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, ".")
from tests.test_dropin import _capture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.path.join(ROOT, "oracle", "_ref")
FREQS = ["136.975", "136.850", "136.725", "136.800", "136.650", "136.775", "136.900", "136.675"]

if __name__ == "__main__":
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    out = []
    with tempfile.TemporaryDirectory() as d:
        fmax = max(float(f) for f in FREQS)
        fos = [int(round((float(f) - fmax) * 1e6)) - 50_000 for f in FREQS]
        # well-formed ACARS-over-AVLC payloads: random ones that happen to look like XID groups send the reference's outxid()
        # (outxid.c:268-299, signed 16-bit group length) into an endless loop, which stops its single consumer thread
        import pathlib
        base, _ = _capture(pathlib.Path(d), fos, nblk=nblk, seed=7, acars=True)
        n = 32768 * nblk * rep
        data = open(base, "rb").read()
        cap = os.path.join(d, "full.cu8")
        with open(cap, "wb") as g:
            for _ in range(rep):
                g.write(data)
        runs = [("file replay, raw cu8, reference blk_thread", "vdlm2dec_file_gpu", {}),
                ("file replay, raw cu8, block pipeline on device", "vdlm2dec_file_gpu_link", {}),
                ("file replay, raw cu8, block pipeline on device, 2^24-sample launches", "vdlm2dec_file_gpu_link", {"VDL2_FILE_BATCH": str(1 << 24)}),
                ("file replay, raw cu8, block pipeline on device, 2^20-sample launches", "vdlm2dec_file_gpu_link", {"VDL2_FILE_BATCH": str(1 << 20)}),
                ("file replay, rtl.c indexing (raw upload, expanded on the device)", "vdlm2dec_file_gpu_link", {"VDL2_RTL_QUIRK": "1"}),
                ("file replay, rtl.c indexing (host expansion to complex float)", "vdlm2dec_file_gpu_link", {"VDL2_RTL_QUIRK": "host"})]
        for nch in (1, 8):
            for name, binary, env in runs:
                b = os.path.join(R, binary)
                if not os.path.exists(b):
                    continue
                t0 = time.perf_counter()
                # no -v: at verbose 2 the reference's own hex dump of XID-looking payloads overruns its 50000-byte text buffer (outxid.c:296 ->
                # out.c:381,396) and the program -- all-reference build included -- dies on this capture's random payloads
                p = subprocess.run([b, "-G", "-E", "-U", "-r", cap, *FREQS[:nch]], env=dict(os.environ, VDL2_FILE_STATS="1", **env), capture_output=True, text=True)
                wall = time.perf_counter() - t0
                m = re.search(r"Replayed (\d+) samples in ([0-9.]+) s", p.stderr)
                rec = {"channels": nch, "rc": p.returncode, "binary": name, "samples": n, "wall_seconds": round(wall, 3), "messages": p.stdout.count("[#")}
                if m:
                    dt = float(m.group(2))
                    rec.update(replay_seconds=dt, stream_msps=round(1e-6 * int(m.group(1)) / dt, 1),
                               channel_msps=round(nch * 1e-6 * int(m.group(1)) / dt, 1), x_real_time=round(int(m.group(1)) / dt / 2e6, 1))
                else:
                    rec["stderr"] = p.stderr[-300:]
                out.append(rec)
                print(json.dumps(rec), flush=True)
        for name, binary in (("reference (rtl.c protocol, wall clock)", "vdlm2dec_cpu"), ("gpu d8psk.o (rtl.c protocol, wall clock)", "vdlm2dec_gpu")):
            b = os.path.join(R, binary)
            if not os.path.exists(b):
                continue
            t0 = time.perf_counter()
            for nch in (1, 8):
              t0 = time.perf_counter()
              p = subprocess.run([b, "-G", "-E", "-U", "-r", "0", *FREQS[:nch]], env=dict(os.environ, VDL2_FAKE_IQ=cap), capture_output=True, text=True)
              wall = time.perf_counter() - t0
              out.append({"channels": nch, "rc": p.returncode, "binary": name, "samples": n, "wall_seconds": round(wall, 3), "messages": p.stdout.count("[#"),
                          "stream_msps_incl_startup": round(1e-6 * n / wall, 1)})
              print(json.dumps(out[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "replay_bench.json"), "w"), indent=1)
