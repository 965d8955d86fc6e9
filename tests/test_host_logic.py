"""CPU tests of the host-side logic and of the kernel's phase 2 SOURCE (vdl2_demod.cuh) compiled
for the host through the fibre warp emulator (tests/emul).  No GPU needed: these catch logic
errors in the lane-parallel restatement (batch trigger search, symbol clock, ballot bit packing,
closed-form de-interleave, lane-per-state header trellis) before any GPU time is spent."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.pyoracle import Oracle
from tests import emul
from tests.parity_util import compare_channel
from vdlm2dec_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/vdl2gpu.h declares."""
    hdr = open(os.path.join(ROOT, "include", "vdl2gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(vdl2_[a-z_0-9]+)\s*\(", hdr)))
    assert set(declared) == set(api.EXPORTS), (declared, api.EXPORTS)
    lib = api.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vdl2_abi_version() == 4


def test_no_gpu_means_loud_failure():
    """No CPU fallback: without a device, create must fail with a message (never silently succeed)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.Vdl2Error, match="no CUDA device|CPU path"):
        api.Vdl2Gpu([(0, 136_975_000, -50_000)])


def test_struct_layouts_match_header():
    assert C.sizeof(api.ChanParam) == 12  # thread_param_t, vdlm2.h:49-52
    assert api.BLOCK_DT.itemsize == 2080 and api.BLOCK_DT.fields["data"][1] == 36
    assert api.SYM_DT.itemsize == 40 and api.SYNC_DT.itemsize == 24 and api.STEP_DT.itemsize == 24
    # row f1 records: vdl2_frame_t (2048 B, hdata at 32) and vdl2_blkstat_t (16 B); the oracle mirrors must agree
    from oracle import pyoracle
    assert api.FRAME_DT.itemsize == 2048 and api.FRAME_DT.fields["hdata"][1] == 32 and api.FRAME_DT == pyoracle.FRAME_DT
    assert api.BLKSTAT_DT.itemsize == 16 and api.BLKSTAT_DT == pyoracle.BLKSTAT_DT
    assert api.AVLC_DT.itemsize == 48 and api.AVLC_DT == pyoracle.AVLC_DT and api.AVLC_DT.fields["txt_off"][1] == 40  # vdl2_avlc_t (row f4)
    hdr = open(os.path.join(ROOT, "include", "vdl2gpu.h")).read()
    assert "uint8_t hdata[2016];" in hdr and "} vdl2_frame_t;" in hdr and "#define VDL2_ABI_VERSION 4" in hdr


@pytest.mark.parametrize("tile", [2688, 84, 84 * 5, 1000])
def test_phase2_source_matches_oracle(tile):
    """Same decimated stream (oracle tap T1) into the emulated warp: blocks bit exact, symbols within
    tolerance, for tile sizes that put every state transition on both sides of a tile edge."""
    n = 1_500_000
    spec = synth.standard_channel(seed=31, nsamples=n, Fo=75_000, period=40_000, payload_bytes=(14, 700))
    iq = synth.render_channel(spec, n)
    o = Oracle("port", Fo=75_000).feed(iq)
    assert len(o.blocks) >= 4
    b, st, sy, sm = emul.demod(o.dumps, tile)
    rep = compare_channel(o, b, sy, sm, None, st)
    assert rep["gi_flips"] <= 2


@pytest.mark.parametrize("tile", [2688, 84 * 12])
@pytest.mark.parametrize("mode", [0x100, 0x200])
def test_phase2_speculative_pass_a(tile, mode):
    """The kernel screens the idle steps of a tile BEFORE it has the previous tile's state, from a guess of the
    tick clock (idle_prepass); idle_run verifies the guess.  With the right guess (0x100) and with a wrong one /
    a stale 'idle' guess during a burst (0x200) the events must equal the plain screened search and the oracle."""
    n = 1_500_000
    spec = synth.standard_channel(seed=33, nsamples=n, Fo=-125_000, period=45_000, payload_bytes=(14, 500))
    iq = synth.render_channel(spec, n)
    o = Oracle("port", Fo=-125_000).feed(iq)
    assert len(o.blocks) >= 4
    b0, _, sy0, sm0 = emul.demod(o.dumps, tile, want_steps=False)
    n_pre = emul.burst_pre_symbols(mode == 0x200)
    b1, _, sy1, sm1 = emul.demod(o.dumps, tile, flags=mode, want_steps=False)
    # tiles that start inside a burst whose header is known also take the burst phases ahead of the chain (BurstPre): on the true
    # symbol grid (0x100) they are consumed, on a wrong grid or tap phase (0x200) they must be ignored
    n_now = emul.burst_pre_symbols(mode == 0x200) - n_pre
    assert n_now % 1_000_000 > 100
    # ... and where the burst ends inside the tile, the idle search of the rest of it ran ahead of the chain as well (IdlePre.pos0)
    assert n_now // 1_000_000 >= 1 or mode == 0x200   # (with the wrong guess only on some tiles: may not occur)
    assert len(b0) == len(b1) and np.array_equal(b0["data"], b1["data"]) and np.array_equal(b0["sync_dump"], b1["sync_dump"])
    assert sy0.tobytes() == sy1.tobytes() and sm0.tobytes() == sm1.tobytes()
    compare_channel(o, b1, sy1, sm1, None, None)


@pytest.mark.parametrize("nlbyte_class", ["le2", "le30", "le67", "gt67", "zero", "rows8"])
def test_phase2_edge_lengths(nlbyte_class):
    length = {"le2": 1992 + 12, "le30": 1992 + 8 * 20, "le67": 1992 + 8 * 50, "gt67": 1992 + 8 * 100, "zero": 1992,
              "rows8": 1992 * 7 + 900}[nlbyte_class]
    rng = np.random.default_rng(6)
    tx = synth.Burst(rng.integers(0, 2, size=length, dtype=np.uint8))
    pidx = synth.burst_phase_indices(tx, rng=rng)
    n = int((len(pidx) + 40) * 2_000_000 / 10500)
    spec = synth.ChannelSpec(-300_000, [dict(burst=tx, phase_idx=pidx, start=2500.0, amp=50.0, cfo=-200.0)], noise_sigma=4.0, seed=3)
    o = Oracle("port", Fo=-300_000).feed(synth.render_channel(spec, n))
    assert len(o.blocks) == 1 and np.array_equal(o.blocks[0]["data"], tx.expected_data)
    b, st, sy, sm = emul.demod(o.dumps, 2688)
    compare_channel(o, b, sy, sm, None, st)


def test_phase2_invalid_header_and_noise():
    rng = np.random.default_rng(8)
    for length in (40, 1992 * 8 + 100):
        tx = synth.Burst(np.ones(64, dtype=np.uint8), length_override=length)
        pidx = synth.burst_phase_indices(tx)
        spec = synth.ChannelSpec(-50_000, [dict(burst=tx, phase_idx=pidx, start=3000.0, amp=60.0)], noise_sigma=3.0, seed=2)
        o = Oracle("port", Fo=-50_000).feed(synth.render_channel(spec, 120_000))
        b, st, sy, sm = emul.demod(o.dumps, 84 * 3)
        compare_channel(o, b, sy, sm, None, st)
    noise = (rng.standard_normal(84_000) + 1j * rng.standard_normal(84_000)).astype(np.complex64)
    b, st, sy, sm = emul.demod(noise, 2688)
    assert len(b) == 0 and len(sy) == 0 and len(st) == 42_000


def test_synth_roundtrip_helpers():
    assert synth.fcs16(b"123456789") == 0x906E  # X.25 check value
    tx = synth.make_burst(np.random.default_rng(1), 200)
    assert tx.valid and tx.tx_bits.size == 25 + 8 * len(tx._byte_order())
    assert synth.scrambler_sequence(8).tolist() == [1, 1, 0, 1, 0, 0, 1, 0] or len(synth.scrambler_sequence(8)) == 8


def test_phase2_noise_free_tail_is_the_only_place_phase_is_ill_conditioned():
    """Found by a randomized sweep of the emulated phase-2 source against the oracle (tools/fuzz_phase2.py, 2 000 cases over Fo, amplitude,
    noise, burst spacing, tile size and speculation mode: every case with noise inside the 1e-5 rad bar; 6 noise-free cases at
    |Fo| = 200 kHz outside it, at most 5.9e-5 rad).  In noise-free input the last symbols of a burst are read after the transmitter stopped:
    what is left is the -0.37 LSB residue of the cu8 conversion, 0.07 LSB after the channel filter against 45+ LSB inside the
    burst, and the phase of a 0.07 LSB vector moves by 3e-5 rad when the filter sum differs in its last bit (the kernel sums
    in pairs, the reference sequentially).  Pinned here: the deviation is confined to such symbols, stays below 1e-4 rad,
    and changes nothing that leaves the demodulator (hard bits, Gray index, blocks)."""
    n, fo = 800_000, 200_000
    spec = synth.standard_channel(seed=1003604752, nsamples=n, Fo=fo, period=66063, payload_bytes=(14, 600),
                                  amp=(45.505150349025584, 45.505150349025584 * 1.5), noise_sigma=0.0)
    o = Oracle("port", Fo=fo).feed(synth.render_channel(spec, n))
    b, _, sy, sm = emul.demod(o.dumps, 1008, want_steps=False)
    osm = o.syms
    assert len(sm) == len(osm) > 2000 and np.array_equal(sm["dump"], osm["dump"])
    dev = np.minimum(np.abs(sm["D"].astype(np.float64) - osm["D"]), 2 * np.pi - np.abs(sm["D"].astype(np.float64) - osm["D"]))
    loud = np.array([np.abs(o.dumps[int(d) - 16:int(d) + 1]).min() for d in osm["dump"]])      # weakest dump in the filter window
    burst_level = np.median(loud)
    off = dev >= 1e-5
    assert 0 < off.sum() <= 4 and dev.max() < 1e-4
    assert (loud[off] < 0.01 * burst_level).all() and (dev[loud > 0.05 * burst_level] < 1e-5).all()
    assert np.array_equal(sm["gi"], osm["gi"]) and np.array_equal(sm["v"] > 0.5, osm["v"] > 0.5)
    assert len(b) == len(o.blocks) and all(np.array_equal(x["data"], y["data"]) for x, y in zip(b, o.blocks))


@pytest.mark.parametrize("fmt", ["cu8", "cs8"])
def test_tensor_core_mixer_tables_and_lane_math_match_oracle(fmt):
    """The int8 tensor-core mixer (mix_rows_mma, vdl2_kernel.cu) replayed lane by lane on the CPU from the product's own host
    tables (vdl2_mma_tables.h): TMA box ring with its wait/release schedule, ldmatrix addressing under the 64B swizzle,
    m16n8k32 fragments, dump masks, accumulator start values, FFMA2/shuffle epilogue.  Its decimated stream must match the
    oracle's T1 tap (d8psk.c:366-381) like the kernel's has to, for full and ragged tiles, with zero protocol errors."""
    from oracle.pyoracle import TAP_DUMPS, Oracle
    from tests import emul
    from vdlm2dec_b200 import synth
    worst = 0.0
    for k, (Fo, nrows) in enumerate([(-50_000, 32), (125_000, 32), (-450_000, 7), (425_000, 32), (300_000, 1)]):
        n = nrows * 2000
        spec = synth.standard_channel(seed=900 + k, nsamples=n, Fo=Fo, period=20_000)
        iq = synth.render_channel(spec, n, fmt=fmt)
        got, rc = emul.mma_mix(iq.view(np.uint8).reshape(nrows, 4000), Fo, cu8=(fmt == "cu8"))
        assert rc == 0, f"{rc} ring protocol errors"
        want = Oracle("port", Fo=Fo, taps=TAP_DUMPS).feed(iq, fmt).dumps
        assert len(want) == nrows * 84
        rms = np.sqrt(np.mean(np.abs(want) ** 2))
        err = np.abs(got[:len(want)] - want).max() / rms
        worst = max(worst, err)
        assert err < 1e-6, f"Fo {Fo} rows {nrows}: decimated stream deviates {err:.2e} of rms"
        assert np.isfinite(got.view(np.float32)).all()   # rows past the end of a short tile (zero-filled boxes) are never read, but stay finite
    assert worst > 0.0
