"""Row f4 of SURVEY.md section 8(f), oracle only so far: the fields the reference derives from a frame before it formats text
or JSON (out.c:517-570, icaoaddr out.c:426-435, outacars.c:214-290).  The independent port (oracle/port/vdl2_avlc_port.c ->
binary record) is pinned against the reference's own out() compiled in place (oracle/ref/ref_out_harness.c -> the -J JSON
line) on synthetic ACARS-over-AVLC frames, and against a committed fixture generated from the reference build."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle
from vdlm2dec_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "avlc_golden.json")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvdl2outref.so")), reason="oracle/_ref not built")


def _frame(payload: bytes) -> bytes:
    """flag + payload + FCS16 + flag: the hdata[0..l) the block pipeline hands to out()."""
    fcs = synth.fcs16(payload) if hasattr(synth, "fcs16") else None
    if fcs is None:
        crc = 0xFFFF
        for b in payload:
            crc ^= b
            for _ in range(8):
                crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
        fcs = crc ^ 0xFFFF
    return bytes([0x7E]) + payload + bytes([fcs & 0xFF, fcs >> 8, 0x7E])


def _cases(n=60, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    labels = ["H1", "10", "Q0", "_\x7f", "SA", "5Z", "B6"]
    for i in range(n):
        text = "".join(chr(int(c)) for c in rng.integers(0x20, 0x7F, int(rng.integers(0, 200))))
        reg = ["G-ABCD", "N123AB", ".F-GXYZ", "9V-SKA", "A9C-KA", "HBJCA", "OOSNA"][i % 7]
        pay = bytearray(synth.acars_frame(int(rng.integers(1, 1 << 24)), reg, labels[i % len(labels)], text,
                                          msgno="M%02dA" % (i % 100), flight="AB%04d" % (i * 7 % 10000), ground=int(rng.integers(1, 1 << 24))))
        out.append(_frame(bytes(pay)))
    return out


def _check(frame: bytes, js: str):
    """Record of the port against the JSON the reference printed for the same frame."""
    r = pyoracle.avlc_extract(frame)
    assert pyoracle.AVLC_KINDS[r["kind"]] == "acars" and js.startswith("{")
    j = json.loads(js)
    if r["fromair"]:
        assert j["icao"] == int(r["faddr"]) & 0xFFFFFF and j["toaddr"] == int(r["taddr"]) & 0xFFFFFF
        assert j["hex"] == "%06X" % (int(r["faddr"]) & 0xFFFFFF)
    else:
        assert j["fromaddr"] == int(r["faddr"]) & 0xFFFFFF and j["icao"] == int(r["taddr"]) & 0xFFFFFF
    assert j.get("is_response", 0) == r["rep"] and bool(j.get("is_onground", 0)) == bool(r["gnd"] and r["fromair"])
    assert j["mode"] == chr(r["mode"]) and j["label"] == bytes(r["label"]).decode("latin-1")
    assert j["block_id"] == chr(r["bid"])
    assert (j["ack"] is False and r["ack"] == 0x15) or j["ack"] == chr(r["ack"])
    assert j["tail"].replace("-", "") == bytes(r["reg"]).decode("latin-1").lstrip(".").replace("-", "")
    if r["mode"] <= ord("Z"):
        assert j["flight"] == bytes(r["fid"][:r["nfid"]]).decode("latin-1") and j["msgno"] == bytes(r["no"][:r["nno"]]).decode("latin-1")
    txt = bytes(b & 0x7F for b in frame[r["txt_off"]:r["txt_off"] + r["txt_len"]]).decode("latin-1")
    assert j.get("text", "") == txt.split("\x00")[0]
    assert ("end" in j) == (r["be"] == 0x17)


@needs_ref
def test_port_matches_reference_json():
    frames = _cases()
    for f in frames:
        _check(f, pyoracle.out_json(f, t=1577836800.0))
    # uplink (ground station is the source), response bit, on-ground bit, block end 0x17, no-text block
    pay = bytearray(synth.acars_frame(0x3C6589, "D-AIMA", "H1", "UPLINK TEXT"))
    pay[0:4], pay[4:8] = pay[4:8], pay[0:4]
    pay[4] |= 2
    pay[0] |= 2
    f = _frame(bytes(pay))
    r = pyoracle.avlc_extract(f)
    assert not r["fromair"] and r["rep"] == 1 and r["gnd"] == 1
    _check(f, pyoracle.out_json(f))


def test_kinds_and_edges():
    base = synth.acars_frame(0x400A0B, "G-ABCD", "H1", "HELLO")
    ok = _frame(base)
    assert pyoracle.AVLC_KINDS[pyoracle.avlc_extract(ok)["kind"]] == "acars"
    bad = bytearray(ok)
    bad[20] ^= 0x04
    assert pyoracle.AVLC_KINDS[pyoracle.avlc_extract(bytes(bad))["kind"]] == "acars_badcrc"
    xid = _frame(base[:9] + bytes([0x82, 0x80, 0x00, 0x01, 0x00]))
    assert pyoracle.AVLC_KINDS[pyoracle.avlc_extract(xid)["kind"]] == "xid"
    empty = _frame(base[:9])                                   # l = 13: addresses + control only
    r = pyoracle.avlc_extract(empty)
    assert pyoracle.AVLC_KINDS[r["kind"]] == "empty" and r["info_len"] == 0 and len(empty) == 13
    other = _frame(base[:9] + b"\x01\x02\x03\x04")
    r = pyoracle.avlc_extract(other)
    assert pyoracle.AVLC_KINDS[r["kind"]] == "other" and r["info_off"] == 10 and r["info_len"] == 4
    r = pyoracle.avlc_extract(ok)
    assert r["faddr"] >> 24 == 1 and r["faddr"] & 0xFFFFFF == 0x400A0B and r["taddr"] >> 24 == 2 and r["lc"] == 0


def test_golden_fixture():
    """Records of the port against the JSON lines the reference build printed when the fixture was made
    (tests/golden/make_avlc_golden.py); runs where /root/reference is absent."""
    gold = json.load(open(GOLD))
    assert len(gold) >= 20
    for g in gold:
        _check(bytes.fromhex(g["frame"]), g["json"])


# ---------------------------------------------------------------- the product's per-frame walk (vdl2_avlc.cuh) against the port
def _frame_records(n_acars=300, n_other=700, seed=21):
    """vdl2_frame_t records: well-formed ACARS frames, the same with damaged bytes, truncated ones, XID-looking, empty, and
    random information fields of random length (1 .. 2000 octets)."""
    rng = np.random.default_rng(seed)
    frames = [bytes(f) for f in _cases(n=n_acars, seed=seed)]
    for i in range(n_other):
        kind = i % 7
        base = bytearray(frames[int(rng.integers(0, n_acars))])
        if kind == 0:                                   # one damaged octet somewhere
            base[int(rng.integers(1, len(base) - 1))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:                                 # truncated (headers of every length, down to nothing)
            base = base[:int(rng.integers(1, len(base)))]
        elif kind == 2:                                 # random information field
            base = bytearray([0x7E]) + bytearray(rng.integers(0, 256, int(rng.integers(9, 2000)), dtype=np.uint8).tobytes()) + b"\x00\x00\x7e"
        elif kind == 3:                                 # XID group
            base = base[:10] + bytes([0x82]) + bytes(rng.integers(0, 256, int(rng.integers(3, 60)), dtype=np.uint8).tobytes()) + b"\x00\x00\x7e"
        elif kind == 4:                                 # no information field
            base = base[:10] + b"\x00\x00\x7e"
        elif kind == 5:                                 # ACARS header, random body (CRC fails; sometimes mode > 'Z' / bid > '9' paths)
            base = base[:13] + bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8).tobytes()) + b"\x7f"
        else:                                           # valid CRC over a random 7-bit body: exercises every branch of the field walk
            body = bytes(rng.integers(0, 128, int(rng.integers(1, 60)), dtype=np.uint8).tobytes())
            crc = 0
            for b in body:
                crc ^= b
                for _ in range(8):
                    crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
            base = base[:13] + body + bytes([crc & 0xFF, crc >> 8, 0x7F]) + b"\x00\x00\x7e"
        frames.append(bytes(base))
    rec = np.zeros(len(frames), pyoracle.FRAME_DT)
    for i, f in enumerate(frames):
        rec[i]["len"] = len(f)
        rec[i]["block"] = i
        rec[i]["hdata"][:len(f)] = np.frombuffer(f, np.uint8)
    return rec, frames


def test_product_walk_equals_port_on_host():
    """vdl2_avlc.cuh compiled for the host (tests/emul/avlc_host.cpp) == the oracle port, byte for byte, on 1000 mixed frames;
    every payload class and both branches of the message-number / flight-id walk occur."""
    from tests import emul
    rec, frames = _frame_records()
    got = emul.avlc(rec)
    want = np.array([pyoracle.avlc_extract(f) for f in frames], dtype=pyoracle.AVLC_DT)
    assert got.tobytes() == want.tobytes()
    kinds = {pyoracle.AVLC_KINDS[k] for k in want["kind"]}
    assert kinds == set(pyoracle.AVLC_KINDS)
    ac = want[want["kind"] == 2]
    assert (ac["nno"] == 4).any() and (ac["nno"] == 0).any() and (ac["txt_len"] == 0).any() and (ac["txt_len"] > 100).any()


@pytest.mark.gpu
def test_avlc_kernel_equals_port_on_gpu():
    """vdl2_avlc_extract through the C ABI: records byte for byte equal to the oracle port."""
    from vdlm2dec_b200 import api
    rec, frames = _frame_records()
    g = api.Vdl2Gpu([(0, 136_975_000, -50_000)], max_samples=200_000)
    got = g.avlc_extract(rec)
    assert len(g.avlc_extract(rec[:0])) == 0
    g.close()
    want = np.array([pyoracle.avlc_extract(f) for f in frames], dtype=pyoracle.AVLC_DT)
    assert got.tobytes() == want.tobytes()


@needs_ref
def test_header_fields_match_reference_text_for_every_address_type():
    """Frames that are NOT ACARS (so no JSON): the addresses, their types, command/response, the on-ground bit and the payload
    class of the port against the TEXT the reference's out() prints (outaddr out.c:437-470, out.c:545-551,570)."""
    import re
    rng = np.random.default_rng(3)
    names = {0: "T0:%06X", 1: "Aircraft:%06X", 2: "T2:%06X", 3: "T3:%06X", 4: "GroundA:%06X", 5: "GroundD:%06X", 6: "T6:%06X"}
    n = 0
    for i in range(400):
        st, dt = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        src = synth.avlc_addr((st << 24) | int(rng.integers(0, 1 << 24)), int(rng.integers(0, 2)), True)
        dst = synth.avlc_addr((dt << 24) | int(rng.integers(0, 1 << 24)), int(rng.integers(0, 2)), False)
        info = bytes(rng.integers(0, 256, int(rng.integers(0, 30)), dtype=np.uint8).tobytes())
        if info and info[0] in (0x82, 0xFF):            # XID / ACARS parsing of random bytes is not what this test is about
            info = b"\x01" + info[1:]
        f = _frame(dst + src + bytes([int(rng.integers(0, 256))]) + info)
        r = pyoracle.avlc_extract(f)
        assert r["faddr"] >> 24 == st and r["taddr"] >> 24 == dt
        text = pyoracle.out_text(f)
        m = re.search(r"\n(Command|Response) from (.*?)\((on ground|airborne)\) to (.*?)\n", text)
        assert m, text
        assert (m.group(1) == "Response") == bool(r["rep"])
        want_from = "All " if st == 7 else names[st] % (int(r["faddr"]) & 0xFFFFFF) + " "
        want_to = "All " if dt == 7 else names[dt] % (int(r["taddr"]) & 0xFFFFFF) + " "
        assert m.group(2) == want_from and m.group(4) == want_to
        assert (m.group(3) == "on ground") == bool(r["fromair"] and r["gnd"])
        assert ("unknown data" in text) == (pyoracle.AVLC_KINDS[r["kind"]] == "other")
        n += pyoracle.AVLC_KINDS[r["kind"]] == "empty"
    assert n > 3
