"""Row f2 of SURVEY.md section 8(f): the raw-capture replay front end (vdlm2dec_b200/csrc/file_shim.c).

The reference declares `initFile` / `runFileSample` (vdlm2.h:110-111) but never defines or calls them; our object defines
them and, through the `initRtl` / `runRtlSample` seam of rtl.o, lets the reference's UNMODIFIED main.c replay a capture:
`vdlm2dec -r capture.cu8 136.975 ...`.  No librtlsdr, no Cbuff, no per-block barriers: raw bytes go to the GPU in large
batches from a ring of page-locked buffers.

Two tiers:
* CPU (`not gpu`): the same two objects linked against a stand-in for libvdl2gpu that answers from the oracle
  (oracle/ref/fake/fake_vdl2gpu.c, test infrastructure) -- checks the HOST logic (argument seam, centre-frequency rule,
  reader ring and batching, the rtl.c:285-292 index quirk, hand-off, capture-time stamps) against the all-reference
  binary fed the same bytes through a fake dongle.
* GPU: the product binaries (real libvdl2gpu.so): identical text to the all-reference binary in quirk mode, oracle block
  counts in raw mode, block pipeline on the device variant identical to the host one.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from tests.test_dropin import AIR_CPU_BIN, ALL, CPU_BIN, _air_capture, _air_expected, _air_run, _capture, _messages, _run
from vdlm2dec_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILE_BIN = os.path.join(ROOT, "oracle", "_ref", "vdlm2dec_file_gpu")            # blocks -> reference blk_thread
FILE_LINK_BIN = os.path.join(ROOT, "oracle", "_ref", "vdlm2dec_file_gpu_link")  # block pipeline on the device (row f1)
HOSTCHECK_BIN = os.path.join(ROOT, "oracle", "_ref", "vdlm2dec_file_hostcheck")  # oracle-backed stand-in, CPU tier only
AIR_FILE_BIN = os.path.join(ROOT, "oracle", "_ref", "vdlm2dec_air_file_gpu")        # -DWITH_AIR: takes the place of air.o
AIR_HOSTCHECK_BIN = os.path.join(ROOT, "oracle", "_ref", "vdlm2dec_air_file_hostcheck")
LINK_HOSTCHECK_BIN = os.path.join(ROOT, "oracle", "_ref", "vdlm2dec_link_hostcheck")            # rtl.c protocol, -DVDL2_SHIM_LINK
FILE_LINK_HOSTCHECK_BIN = os.path.join(ROOT, "oracle", "_ref", "vdlm2dec_file_link_hostcheck")  # replay, -DVDL2_SHIM_LINK
needs_host = pytest.mark.skipif(not (os.path.exists(CPU_BIN) and os.path.exists(HOSTCHECK_BIN)), reason="replay host-check binary not built")
needs_file = pytest.mark.skipif(not (os.path.exists(CPU_BIN) and os.path.exists(FILE_BIN) and os.path.exists(FILE_LINK_BIN)),
                                reason="replay binaries not built")


def _replay(binary, cap, freqs, extra=ALL, **env):
    e = dict(os.environ, **{k: str(v) for k, v in env.items()})
    e.pop("VDL2_FAKE_IQ", None)
    p = subprocess.run([binary, *extra, "-v", "-r", cap, *freqs], env=e, capture_output=True, text=True, timeout=300)
    assert p.returncode == 1, p.stderr[-2000:]  # exit(1) is the reference's normal exit (main.c:246)
    return p.stdout, p.stderr


def _fos(freqs):
    fmax = max(float(f) for f in freqs)  # rtl.c:123-160: the tuner sits 50 kHz above the highest channel when the span allows
    return [int(round((float(f) - fmax) * 1e6)) - 50_000 for f in freqs]


def _expected(cap, fos, fmt="cu8", dtype=np.uint8):
    """Frames the reference algorithm hands to out() for this capture: the single-threaded oracle per channel (fmt
    "rtl_quirk": the stream as rtl.c:285-292 delivers it; else every sample once, in order), its blocks through the block
    pipeline oracle (rs, HDLC un-stuffing, FCS: vdlm2.c:84-161).  A block cut short by the end of the capture or set off
    by a glitch yields no frame, so this can be fewer than the blocks."""
    from oracle.pyoracle import Oracle, link_decode
    iq = np.fromfile(cap, dtype=dtype)
    n = 0
    for fo in fos:
        blocks = Oracle("port", Fo=fo).feed(iq, fmt).blocks
        if len(blocks):
            n += len(link_decode("port", blocks)[0])
    return n


def _bodies(msgs):
    """Message text without the header line (channel, ppm, stamp): what was decoded, not how it was received."""
    return sorted(m.split("\n", 1)[1] if "\n" in m else "" for m in msgs)


# ---------------------------------------------------------------- CPU tier: host logic against the reference binary
@needs_host
@pytest.mark.parametrize("freqs,batch", [(["136.975"], 100_000), (["136.725", "136.975", "136.825"], 1 << 22)])
def test_replay_host_logic_quirk_identical(tmp_path, freqs, batch):
    """VDL2_RTL_QUIRK=1: same text as the all-reference binary fed the same bytes by a (fake) dongle -- the capture is cut
    into several launches (batch rounded up to whole 65536-byte callbacks) and the channel set-up goes through our
    initRtl (same Fc as rtl.c picks, same Fo per channel, same channel numbering)."""
    fos = _fos(freqs)
    cap, nb = _capture(tmp_path, fos, nblk=40)
    ref_out, ref_err = _run(CPU_BIN, cap, freqs, extra=ALL)
    out, err = _replay(HOSTCHECK_BIN, cap, freqs, VDL2_RTL_QUIRK="host", VDL2_FILE_BATCH=batch)   # file_shim.c's own expansion
    dev_out, dev_err = _replay(HOSTCHECK_BIN, cap, freqs, VDL2_RTL_QUIRK=1, VDL2_FILE_BATCH=batch)  # vdl2_process_host_rtl seam
    assert _messages(dev_out) == _messages(out) and "on the device" in dev_err and "on the host" in err
    fc = re.search(r"fakertl: Fc=(\d+)", ref_err).group(1)
    assert f"Set center freq. to {fc}Hz" in err and "rtl.c block indexing" in err
    assert f"Replayed {40 * 32768} samples" in err
    a, b = _messages(ref_out), _messages(out)
    want = _expected(cap, fos, "rtl_quirk")
    assert len(b) == want > 5
    if len(freqs) == 1:
        assert a == b
    else:  # the reference's header trellis is a racy global with several channel threads (see tests/test_dropin.py)
        assert len(set(a) & set(b)) >= want // 2


@needs_host
def test_replay_host_logic_raw_batches_and_tail(tmp_path):
    """Default (raw) mode: every sample once, in order, whatever the batch size -- odd batch lengths leave a
    sub-millisecond tail that is carried into the next launch; a capture that is not a whole number of callbacks keeps
    its tail (the reference would discard the partial read)."""
    fos = _fos(["136.975", "136.850"])
    cap, nb = _capture(tmp_path, fos, nblk=30)
    raw = np.fromfile(cap, dtype=np.uint8)[:-30_000]      # not a multiple of 65536 any more
    raw.tofile(cap)
    want = _expected(cap, fos)
    outs = []
    for batch in (77_777, 1 << 22):
        out, err = _replay(HOSTCHECK_BIN, cap, ["136.975", "136.850"], VDL2_FILE_BATCH=batch, VDL2_FILE_T0=1577836800)
        assert f"Replayed {raw.size // 2} samples" in err and "rtl.c block indexing" not in err
        outs.append(sorted(r.strip() for r in re.split(r"\n(?=\[#)", out) if r.strip()))
    assert len(outs[0]) == want > 5
    assert outs[0] == outs[1]                             # stamps included: capture time, not wall clock
    assert all(" 01/01/2020 00:00:0" in m.split("\n", 1)[0] for m in outs[0])


@needs_host
def test_replay_host_logic_formats_rate_and_errors(tmp_path):
    """Format by extension / VDL2_FILE_FORMAT, centre override, and the reference's error convention."""
    fos = _fos(["136.975"])
    cap, nb = _capture(tmp_path, fos, nblk=30)
    x = (np.fromfile(cap, dtype=np.uint8).astype(np.float32) - 127.37).astype(np.float32)
    cs16 = str(tmp_path / "cap.cs16")
    np.round(x * 64).astype(np.int16).tofile(cs16)
    want = _expected(cs16, fos, "cs16", np.int16)
    out, _ = _replay(HOSTCHECK_BIN, cs16, ["136.975"])
    assert len(_messages(out)) == want > 5
    anon = str(tmp_path / "cap.dat")
    os.rename(cs16, anon)
    out2, _ = _replay(HOSTCHECK_BIN, anon, ["136.975"], VDL2_FILE_FORMAT="cs16")
    assert _messages(out2) == _messages(out)
    # a capture taken at another centre: 137.000 MHz puts the channel at Fo = -25 kHz, nothing decodes, Fc is reported
    out3, err3 = _replay(HOSTCHECK_BIN, anon, ["136.975"], VDL2_FILE_FORMAT="cs16", VDL2_FILE_FC="137.000")
    assert "Set center freq. to 137000000Hz" in err3 and len(_messages(out3)) == 0
    for args, msg in ((["-r", str(tmp_path / "nope.cu8"), "136.975"], "Failed to open capture"),
                      (["-r", anon], "Need a least one frequency"),
                      (["-r", anon, "136.975", "135.000"], "Frequencies too far apart")):
        p = subprocess.run([HOSTCHECK_BIN, *args], capture_output=True, text=True, timeout=60)
        assert p.returncode != 0 and msg in p.stderr and "Unable to init input" in p.stderr  # main.c:209-213


@needs_host
@pytest.mark.skipif(not (os.path.exists(LINK_HOSTCHECK_BIN) and os.path.exists(FILE_LINK_HOSTCHECK_BIN)), reason="link host-check binaries not built")
def test_link_variant_host_logic(tmp_path):
    """-DVDL2_SHIM_LINK objects (no vdlm2.o / rs.o: frames go straight to out() with a msgblk_t the shim fills in): same text
    and -J JSON as the objects that hand blocks to the reference's blk_thread, behind rtl.c and in the replay build, and in the
    replay build with capture-time stamps the output is byte-identical."""
    from tests.test_dropin import HOSTCHECK_BIN as RTL_HOSTCHECK_BIN
    freqs = ["136.975", "136.850"]
    cap, nb = _capture(tmp_path, _fos(freqs), nblk=40, seed=7, acars=True)
    a = _run(RTL_HOSTCHECK_BIN, cap, freqs, extra=ALL)[0]
    b = _run(LINK_HOSTCHECK_BIN, cap, freqs, extra=ALL)[0]
    assert _messages(a) == _messages(b) and len(_messages(a)) > 5
    t0 = dict(VDL2_FILE_T0=1577836800, VDL2_FILE_BATCH=300_000)
    for extra in (ALL, ("-J",)):
        c = _replay(HOSTCHECK_BIN, cap, freqs, extra=extra, **t0)[0]
        d = _replay(FILE_LINK_HOSTCHECK_BIN, cap, freqs, extra=extra, **t0)[0]
        assert sorted(c.split("\n[#")) == sorted(d.split("\n[#")) and len(c) > 1000
    assert '"text":"HELLO VDL2 NUMBER 0' in d


@needs_host
def test_replay_host_logic_eight_channels(tmp_path):
    """MAXNBCHANNELS channels from one stream (BASELINE config 2's shape) through the replay objects: every frame the oracle
    finds, channel numbering and frequencies as main.c registered them."""
    freqs = ["136.975", "136.850", "136.725", "136.800", "136.650", "136.775", "136.900", "136.675"]
    fos = _fos(freqs)
    cap, nb = _capture(tmp_path, fos, nblk=24, seed=9, acars=True)
    out = _replay(HOSTCHECK_BIN, cap, freqs, VDL2_FILE_BATCH=250_000)[0]
    msgs = _messages(out)
    assert len(msgs) == _expected(cap, fos) > 16
    seen = {(int(m[2]), m[m.index("F:") + 2:m.index("F:") + 9]) for m in msgs}       # "[#<chn+1> (F:<MHz> ..."
    # (adjacent 25 kHz channels transmit at the same time in this capture; the boxcar channel filter does not always separate them,
    # for the oracle either -- the count above is the parity statement, this is about numbering)
    assert seen <= {(i + 1, f) for i, f in enumerate(freqs)} and len(seen) >= 6


@needs_host
def test_replay_host_logic_capture_from_a_pipe(tmp_path):
    """`-r -` reads standard input (e.g. from rtl_sdr): same messages as replaying the file."""
    fos = _fos(["136.975"])
    cap, nb = _capture(tmp_path, fos, nblk=30)
    want = _messages(_replay(HOSTCHECK_BIN, cap, ["136.975"], VDL2_FILE_BATCH=200_000)[0])
    with open(cap, "rb") as f:
        p = subprocess.run([HOSTCHECK_BIN, *ALL, "-v", "-r", "-", "136.975"], stdin=f, capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, VDL2_FILE_BATCH="200000"))
    assert p.returncode == 1 and f"Replayed {30 * 32768} samples" in p.stderr
    assert _messages(p.stdout) == want and len(want) > 5


@needs_host
def test_replay_host_logic_interrupt_does_not_hang(tmp_path):
    """Ctrl-C: main.c's handler calls exit() on the main thread, which in the replay build is the thread that holds the
    shim's busy lock while it feeds the GPU; the at-exit quiesce must not wait for itself."""
    import signal
    import time
    fos = _fos(["136.975", "136.850", "136.725"])
    cap, nb = _capture(tmp_path, fos, nblk=200)
    data = open(cap, "rb").read()
    with open(cap, "ab") as f:
        for _ in range(7):
            f.write(data)                                   # ~2 s of work for the oracle-backed stand-in
    p = subprocess.Popen([HOSTCHECK_BIN, "-r", cap, "136.975", "136.850", "136.725"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                         env=dict(os.environ, VDL2_FILE_BATCH=str(1 << 24)))
    time.sleep(0.4)
    p.send_signal(signal.SIGINT)
    try:
        assert p.wait(timeout=20) == 1                      # sighandler: stopVdlm2(); exit(1)  (main.c:106-110)
    finally:
        if p.poll() is None:
            p.kill()


def test_centre_frequency_rule_matches_rtl_c(tmp_path):
    """centre_for() in file_shim.c against the rule of rtl.c:123-160 restated here (first Fc from max+50 kHz downwards,
    1 Hz steps, with every channel between 50 kHz and fs/2-50 kHz away and Fc not the midpoint of two neighbours)."""
    if not os.path.exists(HOSTCHECK_BIN):
        pytest.skip("replay host-check binary not built")
    empty = tmp_path / "e.cu8"
    empty.write_bytes(b"")

    def rule(fd, fs=2_000_000, step=25_000):
        fd = sorted(fd)
        fc = fd[-1] + 2 * step
        while fc > fd[0] - 2 * step:
            if all(2 * step <= abs(fc - f) <= fs // 2 - 2 * step for f in fd) and \
               all(fc - fd[i - 1] != fd[i] - fc for i in range(1, len(fd))):
                break
            fc -= 1
        return fc

    for freqs in (["136.975"], ["136.650", "136.975"], ["136.725", "136.975", "136.775", "136.875"],
                  ["136.000", "136.975"], ["136.000", "137.000", "136.500"], ["136.975", "135.200"]):
        p = subprocess.run([HOSTCHECK_BIN, "-v", "-r", str(empty), *freqs], capture_output=True, text=True, timeout=60)
        fd = [int(1_000_000 * float(f)) for f in freqs]
        assert f"Set center freq. to {rule(fd)}Hz" in p.stderr, (freqs, p.stderr)


def _air_replay(binary, cap, freqs, fs, extra=ALL, **env):
    e = dict(os.environ, VDL2_FILE=cap, VDL2_FILE_RATE=str(fs), **{k: str(v) for k, v in env.items()})
    p = subprocess.run([binary, *extra, "-v", *freqs], env=e, capture_output=True, text=True, timeout=300)
    assert p.returncode == 1, p.stderr[-2000:]
    return p.stdout, p.stderr


@pytest.mark.skipif(not (os.path.exists(AIR_CPU_BIN) and os.path.exists(AIR_HOSTCHECK_BIN)), reason="Airspy replay host-check binary not built")
@pytest.mark.parametrize("fs,freqs", [(6_000_000, ["136.975"]), (5_000_000, ["136.975"]), (5_000_000, ["136.975", "136.725"])])
def test_air_replay_host_logic_identical(tmp_path, fs, freqs):
    """Airspy build (file_shim.c -DWITH_AIR in the place of air.o, capture named by VDL2_FILE): same centre frequency as
    air.c's chooseFc (IF-filter shift at 5 Msps included), same mixer offsets, same text as the all-reference build fed the
    same float32 real samples by a (fake) Airspy -- for one channel; two channel threads race on the reference's header
    trellis, so there the oracle's count and a common majority are demanded."""
    cap, fos = _air_capture(tmp_path, fr_mhz=freqs, fs=fs, nblk=96)
    ref_out, ref_err = _air_run(AIR_CPU_BIN, cap, freqs, fs)
    out, err = _air_replay(AIR_HOSTCHECK_BIN, cap, freqs, fs, VDL2_FILE_BATCH=1_000_003)
    fc = re.search(r"fakeairspy: Fc=(\d+)", ref_err).group(1)
    assert f"Set freq. to {fc} hz" in err and f"Replayed {96 * 32768} samples" in err
    a, b = _messages(ref_out), _messages(out)
    assert len(b) == _air_expected(cap, fos, fs) > 5
    if len(freqs) == 1:
        assert a == b
    else:
        assert len(set(a) & set(b)) >= len(b) // 2
    p = subprocess.run([AIR_HOSTCHECK_BIN, "136.975"], env={k: v for k, v in os.environ.items() if k != "VDL2_FILE"},
                       capture_output=True, text=True, timeout=60)
    assert p.returncode != 0 and "Name the capture to replay in VDL2_FILE" in p.stderr


# ---------------------------------------------------------------- GPU tier: the product binaries
@pytest.mark.gpu
@needs_file
def test_replay_quirk_mode_identical_to_reference(tmp_path):
    """The product binary in VDL2_RTL_QUIRK=1 mode prints what the all-reference binary prints for the same bytes."""
    freqs = ["136.975"]
    cap, nb = _capture(tmp_path, _fos(freqs), nblk=60)
    a = _messages(_run(CPU_BIN, cap, freqs, extra=ALL)[0])
    b = _messages(_replay(FILE_BIN, cap, freqs, VDL2_RTL_QUIRK=1, VDL2_FILE_BATCH=400_000)[0])      # expanded on the device
    c = _messages(_replay(FILE_BIN, cap, freqs, VDL2_RTL_QUIRK="host", VDL2_FILE_BATCH=400_000)[0])  # expanded like rtl.c, on the host
    assert len(b) == _expected(cap, _fos(freqs), "rtl_quirk") > 5 and a == b and a == c


@pytest.mark.gpu
def test_process_host_rtl_quirk_bit_exact_with_oracle():
    """vdl2_process_host_rtl through the C ABI: raw cu8 callbacks in, blocks bit exact with the oracle fed the stream as
    rtl.c:285-292 lays it out, and with the same handle type fed the host-expanded complex floats; in several calls so the
    sub-millisecond tail (768 samples per callback) is carried across them."""
    from oracle.pyoracle import Oracle
    from tests.parity_util import make_channels
    from vdlm2dec_b200 import api
    n = 32768 * 24
    specs, iq = make_channels(1, n, seed=6)
    fo = specs[0].Fo
    want = Oracle("port", Fo=fo).feed(iq[0], "rtl_quirk").blocks
    g = api.Vdl2Gpu([(0, 136_975_000, fo)], fmt="cf32", max_samples=32768 * 8)
    for k in range(0, 24, 8):
        g.process_rtl(iq[0][k * 65536:(k + 8) * 65536])
    got = g.drain_blocks()
    g.close()
    wide = np.zeros((24, 32768, 2), np.float32)
    wide[:, 1:, :] = (iq[0].reshape(24, 32768, 2).astype(np.float32) - np.float32(127.37))[:, :-1, :]
    g2 = api.Vdl2Gpu([(0, 136_975_000, fo)], fmt="cf32", max_samples=n)
    g2.process(wide.reshape(1, -1))
    got2 = g2.drain_blocks()
    g2.close()
    assert len(got) == len(want) > 3 and got.tobytes() == got2.tobytes()
    for a, b in zip(got, want):
        assert a["nbrow"] == b["nbrow"] and a["nlbyte"] == b["nlbyte"] and a["sync_dump"] == b["sync_dump"]
        assert bytes(a["data"]) == bytes(b["data"])
    with pytest.raises(api.Vdl2Error, match="whole"):
        g3 = api.Vdl2Gpu([(0, 136_975_000, fo)], fmt="cf32", max_samples=n)
        g3.process_rtl(iq[0][:2000])


@pytest.mark.gpu
@needs_file
def test_replay_raw_mode_and_device_block_pipeline(tmp_path):
    """Raw cu8 straight to the GPU (2 B/sample, odd batch length): the oracle's block count for the un-quirked stream,
    the same decoded frames as the quirk-mode run wherever both decode, and the f1 variant (RS / HDLC / FCS on the
    device, frames to out()) prints exactly what the variant with the reference's blk_thread prints."""
    freqs = ["136.975", "136.850", "136.725"]
    fos = _fos(freqs)
    cap, nb = _capture(tmp_path, fos, nblk=60, seed=7, acars=True)
    want = _expected(cap, fos)
    t0 = dict(VDL2_FILE_T0=1577836800)
    raw = _replay(FILE_BIN, cap, freqs, VDL2_FILE_BATCH=333_333, **t0)[0]
    one = _replay(FILE_BIN, cap, freqs, **t0)[0]                       # the whole capture in one launch
    link = _replay(FILE_LINK_BIN, cap, freqs, VDL2_FILE_BATCH=333_333, **t0)[0]
    split = lambda t: sorted(r.strip() for r in re.split(r"\n(?=\[#)", t) if r.strip())
    assert len(split(raw)) == want > 5
    assert split(raw) == split(one) == split(link)                     # stamps are capture time: compared too
    quirk = _replay(FILE_BIN, cap, freqs, VDL2_RTL_QUIRK=1)[0]
    common = set(_bodies(_messages(raw))) & set(_bodies(_messages(quirk)))
    assert len(common) >= want * 3 // 4
    ja = _replay(FILE_BIN, cap, freqs, extra=("-J",), **t0)[0]
    jb = _replay(FILE_LINK_BIN, cap, freqs, extra=("-J",), **t0)[0]
    assert sorted(l for l in ja.splitlines() if l.startswith("{")) == sorted(l for l in jb.splitlines() if l.startswith("{"))
    assert '"text":"HELLO VDL2 NUMBER 0' in ja


@pytest.mark.gpu
def test_pinned_host_buffers_through_the_abi():
    """vdl2_host_alloc / vdl2_host_free: a page-locked buffer is usable as vdl2_process_host input and gives the same
    blocks as a pageable one."""
    import ctypes as C
    from tests.parity_util import make_channels
    from vdlm2dec_b200 import api
    lib = api.load_library()
    n = 600_000
    specs, iq = make_channels(1, n, seed=4)
    p = C.c_void_p()
    assert lib.vdl2_host_alloc(2 * n, C.byref(p)) == 0 and p.value
    C.memmove(p.value, iq[0].ctypes.data, 2 * n)
    res = []
    for src in (p.value, iq[0].ctypes.data):
        g = api.Vdl2Gpu([(0, 136_975_000, specs[0].Fo)], max_samples=n)
        assert lib.vdl2_process_host(g.h, C.c_void_p(src), n, 0) == 0
        res.append(g.drain_blocks())
        g.close()
    assert len(res[0]) > 0 and res[0].tobytes() == res[1].tobytes()
    assert lib.vdl2_host_free(p) == 0 and lib.vdl2_host_free(None) == 0


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(AIR_CPU_BIN) and os.path.exists(AIR_FILE_BIN)), reason="Airspy replay binaries not built")
@pytest.mark.parametrize("fs", [6_000_000, 5_000_000])
def test_air_replay_identical_to_reference(tmp_path, fs):
    """The Airspy product binary (float32 real samples, 4 B/sample raw to the GPU): same text as the all-reference build."""
    freqs = ["136.975"]
    cap, fos = _air_capture(tmp_path, fr_mhz=freqs, fs=fs, nblk=96)
    a = _messages(_air_run(AIR_CPU_BIN, cap, freqs, fs)[0])
    b = _messages(_air_replay(AIR_FILE_BIN, cap, freqs, fs, VDL2_FILE_BATCH=1_000_003)[0])
    assert len(b) == _air_expected(cap, fos, fs) > 5 and a == b
