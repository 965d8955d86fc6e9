"""GPU parity tests of the block pipeline (vdl2_link.cu) through the C ABI, against the
CPU oracle (oracle/port/vdl2_link_port.c, itself pinned against the reference's vdlm2.c + rs.c + crc.c):
frames, rs() results per row, corrected rows and consumed byte counts BIT EXACT."""
import os

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import link_decode
from tests.link_util import make_blocks
from tests.parity_util import make_channels
from vdlm2dec_b200.api import Vdl2Gpu

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "link_golden.npz")


def _same(fg, sg, rg, fo, so, ro):
    assert len(fg) == len(fo), (len(fg), len(fo))
    for x, y in zip(fg, fo):
        assert x["block"] == y["block"] and x["len"] == y["len"] and x["chn"] == y["chn"] and x["sync_dump"] == y["sync_dump"]
        assert np.array_equal(x["hdata"][:x["len"]], y["hdata"][:y["len"]])
    assert np.array_equal(sg["rs"], so["rs"]), np.argwhere(sg["rs"] != so["rs"])[:5]
    assert np.array_equal(sg["nframes"], so["nframes"]) and np.array_equal(sg["nbytes"], so["nbytes"])
    if rg is not None:
        assert np.array_equal(rg, ro)


@pytest.fixture(scope="module")
def gpu():
    g = Vdl2Gpu([(0, 136_975_000, -50_000)], max_samples=200_000)
    yield g
    g.close()


@pytest.mark.parametrize("seed,n", [(1, 300), (2, 300), (3, 1500)])
def test_link_kernel_equals_oracle(gpu, seed, n):
    blocks = make_blocks(seed, n)
    fo, so, ro = link_decode("port", blocks)
    fg, sg, rg = gpu.link_decode(blocks)
    _same(fg, sg, rg, fo, so, ro)
    assert len(fg) >= n // 3 and (sg["rs"] < 0).any() and (sg["rs"] > 0).any()


def test_link_kernel_golden_and_edges(gpu):
    g = np.load(GOLDEN)
    blocks = np.frombuffer(g["blocks"].tobytes(), pyoracle.BLOCK_DT).copy()
    f, s, rows = gpu.link_decode(blocks)
    assert np.array_equal(f["len"], g["frame_len"]) and np.array_equal(f["block"], g["frame_block"])
    assert np.array_equal(s["rs"], g["rs"]) and np.array_equal(rows, g["rows_after"])
    assert np.array_equal(np.concatenate([x["hdata"][:x["len"]] for x in f]), g["frame_bytes"])
    # empty call, one block, all-zero block, all-ones block, out-of-range header values (clamped like an index must be)
    f0, s0, _ = gpu.link_decode(blocks[:0])
    assert len(f0) == 0 and len(s0) == 0
    odd = np.zeros(4, pyoracle.BLOCK_DT)
    odd["nbrow"], odd["nlbyte"] = [1, 8, 3, 1], [0, 249, 100, 10]
    odd["data"][1] = 0xFF
    odd["data"][2] = 0x7E
    odd["data"][3, 0, :10] = [0x7E, 0x7E, 1, 2, 3, 4, 5, 6, 7, 0x7E]
    fo, so, ro = link_decode("port", odd)
    fg, sg, rg = gpu.link_decode(odd)
    _same(fg, sg, rg, fo, so, ro)


def test_drain_frames_fused_with_the_demodulator():
    """IQ in -> frames out without the blocks leaving the device in between: same frames as the CPU block
    pipeline applied to the blocks the demodulator produced."""
    nch, n = 6, 1_200_000
    specs, iq = make_channels(nch, n, seed=3)
    chans = [(c, 136_975_000, specs[c].Fo) for c in range(nch)]
    a = Vdl2Gpu(chans, max_samples=n)
    a.process(iq)
    blocks_ref = a.drain_blocks()
    b = Vdl2Gpu(chans, max_samples=n)
    b.process(iq)
    frames, blocks = b.drain_frames()
    assert len(blocks) == len(blocks_ref) >= nch and np.array_equal(blocks["data"], blocks_ref["data"])
    fo, so, _ = link_decode("port", blocks)
    assert len(frames) == len(fo) >= nch
    for x, y in zip(frames, fo):
        assert x["block"] == y["block"] and x["len"] == y["len"] and np.array_equal(x["hdata"][:x["len"]], y["hdata"][:y["len"]])
    st = b.stats()
    assert st["link_launches"] == 1 and st["frames_out"] == len(frames)
    f2, b2 = b.drain_frames()
    assert len(f2) == 0 and len(b2) == 0


def test_link_full_size_round_trip(gpu):
    """Size-independent property at bench scale: 16384 bursts, transmitted frame -> RS-protected rows -> up to the
    code's capacity of byte errors per row -> the kernel must hand back exactly the transmitted frame for every block
    (and agree with the oracle on everything else)."""
    from vdlm2dec_b200 import synth
    from tests.link_util import block_from_burst
    rng = np.random.default_rng(123)
    base, want = [], []
    for i in range(256):
        payload = synth.random_payload(rng, int(rng.integers(14, 1200)))
        b = synth.Burst(synth.hdlc_bits(payload))
        blk = block_from_burst(b, chn=i % 8, sync_dump=i)
        d = blk["data"]
        for r in range(b.nbrow):
            last = r == b.nbrow - 1
            cap = 3 if not last or b.nlbyte > 67 else (2 if b.nlbyte > 30 else (1 if b.nlbyte > 2 else 0))
            width = b.nlbyte if last else 249
            for c in rng.choice(width, size=min(cap, width), replace=False):
                d[r, c] ^= np.uint8(rng.integers(1, 256))
        blk["data"] = d
        base.append(blk)
        fcs = synth.fcs16(payload)
        want.append(bytes([0x7E]) + payload + bytes([fcs & 0xFF, fcs >> 8, 0x7E]))
    blocks = np.tile(np.array(base), 64)
    blocks["sync_dump"] = np.arange(len(blocks))
    f, s, _ = gpu.link_decode(blocks, want_rows=False)
    assert len(f) == len(blocks) == 16384 and np.array_equal(f["block"], np.arange(len(blocks)))
    assert (s["nframes"] == 1).all() and (s["rs"] >= 0).all()
    for i in range(0, len(f), 97):
        assert bytes(f[i]["hdata"][:f[i]["len"]]) == want[i % 256]
    assert np.array_equal(f["len"], np.array([len(w) for w in want] * 64))


def test_drain_frames_packed_equals_records_and_port(tmp_path):
    """Rows f1 + f4 end to end (vdl2_drain_frames_packed): ACARS-over-AVLC bursts on 4 channels of one stream -> the frames leave
    the device ordered (completion order) and packed, with their field records.  Bytes must equal the fixed-record drain of a
    second handle, records must equal the port's walk of the same frames, the order must be (end of burst, channel, length)."""
    from oracle import pyoracle
    from tests.test_dropin import _capture
    fos = [-50_000, -175_000, -300_000, 125_000]
    cap, nb = _capture(tmp_path, fos, nblk=40, seed=9, acars=True)
    iq = np.fromfile(cap, dtype=np.uint8)[None, :]
    n = iq.shape[1] // 2
    chans = [(c, 136_975_000 + fo + 50_000, fo) for c, fo in enumerate(fos)]
    a = Vdl2Gpu(chans, ch_per_stream=len(fos), max_samples=n)
    a.process(iq)
    frames, blocks = a.drain_frames()
    b = Vdl2Gpu(chans, ch_per_stream=len(fos), max_samples=n)
    b.process(iq)
    hdrs, data, recs = b.drain_frames_packed()
    assert len(hdrs) == len(frames) >= nb // 2 > 4
    end = hdrs["sync_dump"] + hdrs["dur"]
    key = list(zip(end.tolist(), hdrs["chn"].tolist(), hdrs["len"].tolist()))
    assert key == sorted(key), "packed frames are not in completion order"
    assert (hdrs["offset"] % 16 == 0).all() and len(data) == int(((hdrs["len"] + 15) // 16 * 16).sum())
    # same frames as the fixed-record drain (which orders by trigger time)
    want = {(int(f["sync_dump"]), int(f["chn"]), int(f["len"])): bytes(f["hdata"][:f["len"]]) for f in frames}
    durs = {(int(x["sync_dump"]), int(x["chn"])): int(x["end_dump"] - x["sync_dump"]) for x in blocks}
    for h in hdrs:
        k = (int(h["sync_dump"]), int(h["chn"]), int(h["len"]))
        assert bytes(data[h["offset"]:h["offset"] + h["len"]]) == want[k]
        assert durs[k[:2]] == h["dur"]
    # field records: the port's walk of the same frames, in the same order
    port = np.array([pyoracle.avlc_extract(bytes(data[h["offset"]:h["offset"] + h["len"]])) for h in hdrs], dtype=pyoracle.AVLC_DT)
    assert recs.tobytes() == port.tobytes()
    assert (recs["kind"] == 2).sum() >= len(hdrs) // 2      # well-formed ACARS bodies: the lane-parallel CRC says good
    assert b.last_pack_ms > 0 and len(b.drain_frames_packed()[0]) == 0
