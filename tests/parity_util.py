"""Shared helpers for the parity tests: run the CPU oracle and the CUDA path on the same
seeded input and compare the tap points T1..T6."""
from __future__ import annotations

import numpy as np

from oracle import pyoracle
from oracle.pyoracle import Oracle
from vdlm2dec_b200 import synth

D_TOL = 1e-5  # rad, absolute: the north_star's soft-symbol tolerance (DESIGN.md "numerics")


def oracle_kind() -> str:
    """The checker the parity tests compare against: the REFERENCE's own d8psk.c + viterbi.c compiled in place (oracle/_ref,
    strict -O2; built here, travels to the GPU box) whenever it is present and loads, else the independent port.  The port is
    pinned bit for bit against the reference in tests/test_oracle.py (CPU tier) and again on the GPU box
    (tests/test_gpu_parity.py::test_port_equals_reference_on_this_box)."""
    global _KIND
    if _KIND is None:
        _KIND = "port"
        if pyoracle.available("ref"):
            try:
                pyoracle.load("ref")
                _KIND = "ref"
            except OSError:
                pass
    return _KIND


_KIND = None


def oracle(chn=0, Fr=136_975_000, Fo=-50_000, fs=2_000_000, sdrclk=500, kind=None, **kw) -> Oracle:
    return Oracle(kind or oracle_kind(), chn=chn, Fr=Fr, Fo=Fo, fs=fs, sdrclk=sdrclk, **kw)


def make_channels(nch, nsamples, seed=1, fs=2_000_000, fmt="cu8", period=60000, fos=None, **kw):
    """nch independent streams; Fo cycles over the legal 25 kHz raster (|Fo| >= 50 kHz)."""
    fos = fos or [f for f in range(-450_000, 475_000, 125_000) if abs(f) >= 50_000]
    specs, iqs = [], []
    for c in range(nch):
        spec = synth.standard_channel(seed=seed * 1000 + c, nsamples=nsamples, Fo=fos[c % len(fos)], fs=fs,
                                      period=period, **kw)
        specs.append(spec)
        iqs.append(synth.render_channel(spec, nsamples, fs=fs, fmt=fmt))
    return specs, np.stack(iqs)


def run_oracle(iq_row, Fo, fmt="cu8", kind=None, fs=2_000_000, sdrclk=500, chn=0, Fr=136_975_000):
    o = Oracle(kind or oracle_kind(), chn=chn, Fr=Fr, Fo=Fo, fs=fs, sdrclk=sdrclk, real_input=(fmt == "f32real"))
    o.feed(iq_row, fmt)
    return o


def wrap_diff(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return np.minimum(d, np.abs(2 * np.pi - d))


def compare_channel(o: Oracle, g_blocks, g_syncs=None, g_syms=None, g_dumps=None, g_steps=None, ndump_limit=None):
    """Returns a dict of findings; raises AssertionError on a parity failure."""
    rep = {}
    ob = o.blocks
    if ndump_limit is not None:
        ob = ob[ob["end_dump"] < ndump_limit]
    rep["blocks"] = (len(ob), len(g_blocks))
    assert len(ob) == len(g_blocks), f"block count oracle {len(ob)} vs gpu {len(g_blocks)}"
    for i, (a, b) in enumerate(zip(ob, g_blocks)):
        for f in ("sync_dump", "end_dump", "nbrow", "nlbyte"):
            assert a[f] == b[f], f"block {i} field {f}: oracle {a[f]} gpu {b[f]}"
        assert np.array_equal(a["data"], b["data"]), f"block {i}: data bytes differ ({(a['data'] != b['data']).sum()})"
        assert abs(float(a["ppm"]) - float(b["ppm"])) <= 1e-4 * max(1.0, abs(float(a["ppm"]))), f"block {i} ppm"
    if g_dumps is not None:
        od = o.dumps[:len(g_dumps)]
        scale = np.sqrt(np.mean(np.abs(od) ** 2)) + 1e-30
        err = np.abs(od - g_dumps).max() / scale
        rep["dumps_relerr"] = float(err)
        assert err < 1e-5, f"decimated stream deviates {err:.3e} of rms"
    if g_syncs is not None:
        os_ = o.syncs
        if ndump_limit is not None:
            os_ = os_[os_["dump"] < ndump_limit]
        assert len(os_) == len(g_syncs), f"sync count oracle {len(os_)} gpu {len(g_syncs)}"
        assert np.array_equal(os_["dump"], g_syncs["dump"]), "sync positions differ"
        assert np.array_equal(os_["clk"], g_syncs["clk"]), "sync timing (clk) differs"
        rep["sync_df_err"] = float(np.abs(os_["df"] - g_syncs["df"]).max()) if len(os_) else 0.0
        assert rep["sync_df_err"] < D_TOL
    if g_syms is not None:
        osy = o.syms
        if ndump_limit is not None:
            osy = osy[osy["dump"] < ndump_limit]
        assert len(osy) == len(g_syms), f"symbol count oracle {len(osy)} gpu {len(g_syms)}"
        if len(osy):
            assert np.array_equal(osy["dump"], g_syms["dump"]), "symbol positions differ"
            rep["sym_D_err"] = float(wrap_diff(osy["D"], g_syms["D"]).max())
            flips = osy["gi"] != g_syms["gi"]
            rep["gi_flips"] = int(flips.sum())
            assert rep["sym_D_err"] < D_TOL, f"soft symbol deviates {rep['sym_D_err']:.3e} rad"
            # The table index is roundf(128*D/pi+128) (d8psk.c:213): a 1e-6 rad difference in D moves a
            # symbol sitting on a rounding boundary to the NEIGHBOURING entry.  That is allowed only
            # there (both D within tolerance of the same boundary) and only as a +-1 step; the hard
            # decisions (what reaches vdlm2.c) must be identical everywhere.
            if flips.any():
                assert np.abs(osy["gi"][flips] - g_syms["gi"][flips]).max() == 1, "Gray index differs by more than one step"
                x = 128.0 * osy["D"][flips].astype(np.float64) / np.pi + 128.0
                assert np.abs(np.abs(x - np.floor(x)) - 0.5).max() < 128 / np.pi * D_TOL, "index flip away from a rounding boundary"
                assert rep["gi_flips"] <= max(1, len(osy) // 500), f"{rep['gi_flips']} Gray index flips in {len(osy)} symbols"
            same = ~flips
            assert np.array_equal(osy["v"][same].view(np.uint32), g_syms["v"][same].view(np.uint32)), "soft bits differ"
            assert np.array_equal(osy["v"] > 0.5, g_syms["v"] > 0.5), "hard bit decisions differ"
    if g_steps is not None:
        ost = o.steps
        if ndump_limit is not None:
            ost = ost[ost["dump"] < ndump_limit]
        assert len(ost) == len(g_steps), f"step count oracle {len(ost)} gpu {len(g_steps)}"
        if len(ost):
            assert np.array_equal(ost["dump"], g_steps["dump"]), "step positions differ"
            rep["step_P_err"] = float(wrap_diff(ost["P"], g_steps["P"]).max())
            m = ost["err"] >= 0
            rel = np.abs(ost["err"][m] - g_steps["err"][m]) / np.maximum(1.0, np.abs(ost["err"][m]))
            rep["step_err_rel"] = float(rel.max()) if m.any() else 0.0
    return rep
