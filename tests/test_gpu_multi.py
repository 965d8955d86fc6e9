"""GPU tests of the ingest and sharding entry points added in ABI version 4 (include/vdl2gpu.h): asynchronous submit with
the page-locked ring, the non-blocking pending-block poll, and several devices behind one handle (SURVEY.md section 8(e):
stream s on device s mod N, no collective, host merge in the order one device would produce)."""
import numpy as np
import pytest

from tests.parity_util import compare_channel, make_channels, run_oracle
from vdlm2dec_b200.api import Vdl2Gpu, Vdl2Multi

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def test_async_submit_copy_equals_synchronous_path():
    """The drop-in's round: vdl2_submit_copy() of a 32768-sample callback buffer that is OVERWRITTEN right after the call
    (like Cbuff, rtl.c:283-294), blocks collected only when vdl2_pending_blocks() reports some, the rest at the end."""
    nch, nblk = 3, 48
    n = 32768 * nblk
    specs, iq = make_channels(nch, n, seed=41, period=90_000)
    chans = [(c, 136_975_000, specs[c].Fo) for c in range(nch)]
    a = Vdl2Gpu(chans, max_samples=n)
    a.process(iq)
    want = a.drain_blocks()
    b = Vdl2Gpu(chans, max_samples=32768)
    buf = np.empty((nch, 65536), np.uint8)
    got, polls_with_blocks = [], 0
    for k in range(nblk):
        buf[:] = iq[:, k * 65536:(k + 1) * 65536]
        b.submit_copy(buf)
        buf[:] = 0xAA        # the caller's buffer is free again
        if b.pending_blocks() > 0:
            polls_with_blocks += 1
            got.append(b.drain_blocks())
    got.append(b.drain_blocks())   # synchronises: everything still in flight
    got = np.concatenate(got)
    got = got[np.lexsort((got["chn"], got["sync_dump"]))]
    assert len(want) >= nch and polls_with_blocks >= 1
    assert want.tobytes() == got.tobytes()
    assert b.stats()["kernel_launches"] == nblk and b.pending_blocks() == 0


@pytest.mark.parametrize("ndev", [2, 3])
def test_multi_handle_union_equals_one_device(ndev):
    """64 channels on one device, and sharded s mod N over N handles (the same ordinal N times when the box has fewer GPUs:
    the host-side split and merge are what is under test here; tests with distinct devices follow): merged blocks bit identical."""
    nch, n = 64, 600_000
    specs, iq = make_channels(nch, n, seed=43, period=70_000)
    chans = [(c, 136_975_000, specs[c].Fo) for c in range(nch)]
    a = Vdl2Gpu(chans, max_samples=n)
    a.process(iq)
    want = a.drain_blocks()
    have = _ndev()
    devices = [d % have for d in range(ndev)]
    m = Vdl2Multi(chans, devices, max_samples=n)
    assert m.ndev == ndev
    m.process(iq)
    got = m.drain_blocks()
    assert len(want) >= nch and want.tobytes() == got.tobytes()
    # and the union really is the reference's answer
    for c in (0, 17, 63):
        compare_channel(run_oracle(iq[c], specs[c].Fo, chn=c), got[got["chn"] == c], ndump_limit=n // 2000 * 84)


def test_multi_handle_on_every_gpu_of_the_box():
    """SURVEY.md section 4.5 on hardware: the union of the shards of N real devices equals the 1-GPU output."""
    have = _ndev()
    if have < 2:
        pytest.skip("one GPU on this box")
    nch, n = 64, 600_000
    specs, iq = make_channels(nch, n, seed=44, period=70_000)
    chans = [(c, 136_975_000, specs[c].Fo) for c in range(nch)]
    a = Vdl2Gpu(chans, max_samples=n)
    a.process(iq)
    want = a.drain_blocks()
    for N in sorted({2, have}):
        m = Vdl2Multi(chans, list(range(N)), max_samples=n)
        m.process(iq)
        got = m.drain_blocks()
        assert len(want) >= nch and want.tobytes() == got.tobytes(), f"{N} devices"
        m.close()


def test_multi_shared_streams_and_ragged_split():
    """8 channels per stream, 5 streams over 3 handles (2 + 2 + 1 streams): channels follow their stream."""
    cps, nstreams, n = 4, 5, 500_000
    fos = [-450_000, -200_000, 175_000, 425_000]
    rng = np.random.default_rng(5)
    from vdlm2dec_b200 import synth
    iq = []
    for s in range(nstreams):
        x = np.zeros(n, dtype=np.complex128)
        for k, fo in enumerate(fos):
            spec = synth.standard_channel(seed=700 + 10 * s + k, nsamples=n, Fo=fo, period=60_000, amp=(14.0, 20.0), noise_sigma=0.0)
            x += synth.render_channel(spec, n, fmt="cf32").astype(np.float64).view(np.complex128)
        x += 3.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        iq.append(synth.quantise(x, "cu8"))
    iq = np.stack(iq)
    chans = [(s * cps + k, 136_000_000 + fos[k], fos[k]) for s in range(nstreams) for k in range(cps)]
    a = Vdl2Gpu(chans, ch_per_stream=cps, max_samples=n)
    a.process(iq)
    want = a.drain_blocks()
    have = _ndev()
    m = Vdl2Multi(chans, [d % have for d in range(3)], ch_per_stream=cps, max_samples=n)
    m.process(iq)
    got = m.drain_blocks()
    assert len(want) >= nstreams * cps and want.tobytes() == got.tobytes()


@pytest.mark.parametrize("fmt,cps", [("cu8", 8), ("cs8", 8), ("cu8", 1), ("cs8", 1)])
def test_channeliser_one_pass_equals_fused_kernel_and_oracle(fmt, cps):
    """Row f3 (vdl2_channelise_device): ONE pass over each shared stream yields the decimated streams of all its channels.
    They must equal the fused kernel's own T1 tap bit for bit (same integer sums) and the oracle's dumps within the T1 bar;
    3 streams x 8 channels, 70 rows (a ragged last tile)."""
    import torch
    from oracle.pyoracle import TAP_DUMPS as O_DUMPS
    from tests.parity_util import oracle
    from vdlm2dec_b200 import synth
    from vdlm2dec_b200.api import TAP_DUMPS
    nstreams, rows = 3, 70          # cps = 1: the kernel's quad-of-dumps store path (one channel per stream)
    n = rows * 2000
    fos = [-450_000, -325_000, -200_000, -75_000, 50_000, 175_000, 300_000, 425_000][:cps] if cps > 1 else [175_000]
    rng = np.random.default_rng(8)
    iq = []
    for s in range(nstreams):
        x = 3.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        for k, fo in enumerate(fos):
            spec = synth.standard_channel(seed=800 + 10 * s + k, nsamples=n, Fo=fo, period=40_000, amp=(12.0, 16.0), noise_sigma=0.0)
            x += synth.render_channel(spec, n, fmt="cf32").astype(np.float64).view(np.complex128)
        iq.append(synth.quantise(x, fmt))
    iq = np.stack(iq)
    chans = [(s * cps + k, 136_000_000 + fos[k] % 1_000_000, fos[k]) for s in range(nstreams) for k in range(cps)]
    g = Vdl2Gpu(chans, fmt=fmt, ch_per_stream=cps, taps=TAP_DUMPS, max_samples=n)
    t = torch.from_numpy(iq.view(np.uint8)).cuda()
    out = torch.zeros((len(chans), rows * 84 + 4), dtype=torch.complex64, device="cuda")
    g.channelise_device(t.data_ptr(), n, t.stride(0), out.data_ptr(), out.stride(0))
    g.sync()
    assert g.stats()["last_kernel_ms"] > 0
    got = out.cpu().numpy()
    assert (got[:, rows * 84:] == 0).all()             # nothing written past the last row
    g.process(iq)                                       # the fused kernel on the same samples, with its T1 tap
    for c, (chn, Fr, fo) in enumerate(chans):
        tap = g.read_dumps(c)
        assert len(tap) == rows * 84 and tap.view(np.uint32).tolist() == got[c, :rows * 84].view(np.uint32).tolist(), f"channel {c}"
        if c % 5 == 0:
            want = oracle(chn=chn, Fr=Fr, Fo=fo, taps=O_DUMPS).feed(iq[c // cps], fmt).dumps
            rms = np.sqrt(np.mean(np.abs(want) ** 2))
            assert np.abs(want - got[c, :rows * 84]).max() < 1e-5 * rms


def test_create_failure_is_reported_and_leaks_nothing():
    """vdl2_create() that fails half way (here: the staging buffer of an absurd max_samples does not fit the device) must
    return the reason through vdl2_last_error(NULL) and tear the partial handle down: the next create succeeds and the
    device's free memory is back where it was."""
    import torch
    from vdlm2dec_b200.api import Vdl2Error
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(3):
        with pytest.raises(Vdl2Error) as ei:
            Vdl2Gpu([(c, 136_975_000, -50_000) for c in range(64)], max_samples=1 << 36)
        assert "failed" in str(ei.value) or "memory" in str(ei.value).lower()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, f"{(free0 - free1) >> 20} MiB still allocated after three failed creates"
    g = Vdl2Gpu([(0, 136_975_000, -50_000)], max_samples=100_000)
    g.process(np.full((1, 2 * 4000), 127, np.uint8))
    assert g.stats()["samples_done"] == 4000
