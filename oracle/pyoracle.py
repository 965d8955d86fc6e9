"""ctypes front end of the CPU checkers in oracle/ -- TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs
(cpu_baseline, --impl reference).  The product package never imports this module.

  Oracle("port")       -> oracle/libvdl2port.so          plain-C restatement
  Oracle("ref")        -> oracle/_ref/libvdl2ref_O2.so   reference d8psk.c compiled in place, strict -O2
  Oracle("ref_fast")   -> oracle/_ref/libvdl2ref_fast.so   same, project flags -Ofast (timing)
  Oracle("ref_native") -> oracle/_ref/libvdl2ref_native.so same, -Ofast -march=native of the build host
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TAP_DUMPS, TAP_STEPS, TAP_SYNCS, TAP_SYMS, TAP_BLOCKS = 1, 2, 4, 8, 16
TAP_ALL = 31

STEP_DT = np.dtype([("dump", "<i8"), ("P", "<f4"), ("err", "<f4"), ("fr", "<f4"), ("pad", "<i4")])
SYNC_DT = np.dtype([("dump", "<i8"), ("clk", "<i4"), ("df", "<f4"), ("ppm", "<f4"), ("P1", "<f4")])
SYM_DT = np.dtype([("dump", "<i8"), ("D", "<f4"), ("P", "<f4"), ("gi", "<i4"), ("v", "<f4", (3,)),
                   ("state_after", "<i4"), ("pad", "<i4")])
BLOCK_DT = np.dtype([("sync_dump", "<i8"), ("end_dump", "<i8"), ("chn", "<i4"), ("Fr", "<i4"), ("ppm", "<f4"),
                     ("nbrow", "<i4"), ("nlbyte", "<i4"), ("data", "u1", (8, 255)), ("pad", "u1", (4,))])
assert STEP_DT.itemsize == 24 and SYNC_DT.itemsize == 24 and SYM_DT.itemsize == 40 and BLOCK_DT.itemsize == 2080

_PATHS = {
    "port": os.path.join(HERE, "libvdl2port.so"),
    "ref": os.path.join(HERE, "_ref", "libvdl2ref_O2.so"),
    "ref_fast": os.path.join(HERE, "_ref", "libvdl2ref_fast.so"),
    "ref_native": os.path.join(HERE, "_ref", "libvdl2ref_native.so"),
}
_LIBS: dict = {}


def build(which: str = "all") -> None:
    """Compile the checkers (port always; oracle/_ref only when /root/reference is mounted)."""
    subprocess.run(["make", "-s", "-C", HERE, which], check=True)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


def load(kind: str):
    if kind in _LIBS:
        return _LIBS[kind]
    path = _PATHS[kind]
    if not os.path.exists(path):
        raise FileNotFoundError(f"oracle library {path} missing (run `make -C oracle`)")
    lib = C.CDLL(path)
    lib.orc_open.restype = C.c_void_p
    lib.orc_open.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint32]
    lib.orc_close.argtypes = [C.c_void_p]
    for f in ("orc_feed_cf32", "orc_feed_f32real", "orc_feed_cs8", "orc_feed_cs16"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.orc_feed_cu8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float]
    lib.orc_feed_rtl_block_quirk.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_tap.restype = C.c_void_p
    lib.orc_tap.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
    lib.orc_clear_taps.argtypes = [C.c_void_p]
    lib.orc_ndump.restype = C.c_int64
    lib.orc_ndump.argtypes = [C.c_void_p]
    lib.orc_kind.restype = C.c_char_p
    lib.orc_table.restype = C.POINTER(C.c_float)
    lib.orc_table.argtypes = [C.c_int]
    lib.orc_nco.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_time_cu8.restype = C.c_double
    lib.orc_time_cu8.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_size_t, C.c_int]
    _LIBS[kind] = lib
    return lib


def table(kind: str, which: int) -> np.ndarray:
    """0 sync word, 1 interpolating filter (65), 2..4 soft demap tables (257)."""
    n = {0: 17, 1: 65}.get(which, 257)
    p = load(kind).orc_table(which)
    return np.ctypeslib.as_array(p, shape=(n,)).copy()


class Oracle:
    """One channel of the CPU path: feed samples, read the taps."""

    def __init__(self, kind: str = "port", chn: int = 0, Fr: int = 136_975_000, Fo: int = -50_000,
                 fs: int = 2_000_000, sdrclk: int = 500, real_input: bool = False, taps: int = TAP_ALL):
        self.lib = load(kind)
        self.kind = kind
        self.h = self.lib.orc_open(chn, Fr, Fo, fs, sdrclk, int(real_input), taps)

    def close(self):
        if self.h:
            self.lib.orc_close(self.h)
            self.h = None

    __del__ = close

    def feed(self, iq: np.ndarray, fmt: str = "cu8"):
        iq = np.ascontiguousarray(iq)
        p = iq.ctypes.data_as(C.c_void_p)
        if fmt == "cu8":
            assert iq.dtype == np.uint8
            self.lib.orc_feed_cu8(self.h, p, iq.size // 2, C.c_float(np.float32(127.37)))
        elif fmt == "cs8":
            assert iq.dtype == np.int8
            self.lib.orc_feed_cs8(self.h, p, iq.size // 2)
        elif fmt == "cs16":
            assert iq.dtype == np.int16
            self.lib.orc_feed_cs16(self.h, p, iq.size // 2)
        elif fmt == "cf32":
            assert iq.dtype == np.float32
            self.lib.orc_feed_cf32(self.h, p, iq.size // 2)
        elif fmt == "f32real":
            assert iq.dtype == np.float32
            self.lib.orc_feed_f32real(self.h, p, iq.size)
        elif fmt == "rtl_quirk":
            assert iq.dtype == np.uint8 and iq.size % 65536 == 0
            for k in range(iq.size // 65536):
                blk = iq[k * 65536:(k + 1) * 65536]
                self.lib.orc_feed_rtl_block_quirk(self.h, blk.ctypes.data_as(C.c_void_p))
        else:
            raise ValueError(fmt)
        return self

    def _tap(self, which, dt):
        n = C.c_size_t(0)
        p = self.lib.orc_tap(self.h, which, C.byref(n))
        if not p or n.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n.value * np.dtype(dt).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt).copy()

    @property
    def nco(self) -> np.ndarray:
        """The channel's oscillator table wf[] (d8psk.c:353-357) as complex64."""
        out = np.zeros(2 * 4096, np.float32)
        n = self.lib.orc_nco(self.h, out.ctypes.data_as(C.c_void_p), 4096)
        return out[:2 * n].view(np.complex64).copy()

    @property
    def dumps(self):
        return self._tap(TAP_DUMPS, np.dtype("<c8"))

    @property
    def steps(self):
        return self._tap(TAP_STEPS, STEP_DT)

    @property
    def syncs(self):
        return self._tap(TAP_SYNCS, SYNC_DT)

    @property
    def syms(self):
        return self._tap(TAP_SYMS, SYM_DT)

    @property
    def blocks(self):
        return self._tap(TAP_BLOCKS, BLOCK_DT)

    @property
    def ndump(self):
        return int(self.lib.orc_ndump(self.h))


FRAME_DT = np.dtype([("block", "<i4"), ("len", "<i4"), ("chn", "<i4"), ("Fr", "<i4"), ("ppm", "<f4"), ("pad", "<i4"),
                     ("sync_dump", "<i8"), ("hdata", "u1", (2016,))])
BLKSTAT_DT = np.dtype([("rs", "i1", (8,)), ("nbytes", "<i4"), ("nframes", "<i4")])
assert FRAME_DT.itemsize == 2048 and BLKSTAT_DT.itemsize == 16
_LINK_PATHS = {"port": os.path.join(HERE, "libvdl2linkport.so"), "ref": os.path.join(HERE, "_ref", "libvdl2linkref.so")}
_LINK_LIBS: dict = {}


def link_available(kind: str) -> bool:
    return os.path.exists(_LINK_PATHS[kind])


def _link_lib(kind: str):
    if kind not in _LINK_LIBS:
        path = _LINK_PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"oracle library {path} missing (run `make -C oracle`)")
        lib = C.CDLL(path)
        lib.orc_link_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        lib.orc_link_time.restype = C.c_double
        lib.orc_link_time.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _LINK_LIBS[kind] = lib
    return _LINK_LIBS[kind]


def link_decode(kind: str, blocks: np.ndarray):
    """Block pipeline behind the demodulator (blk_thread, vdlm2.c:84-161) on the CPU: returns
    (frames, per-block stats, data rows after rs())."""
    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DT)
    n = len(blocks)
    frames = np.zeros(max(4 * n, 16), FRAME_DT)
    stats = np.zeros(n, BLKSTAT_DT)
    rows = np.zeros((n, 8, 255), np.uint8)
    nf = C.c_int(0)
    rc = _link_lib(kind).orc_link_decode(blocks.ctypes.data_as(C.c_void_p), n, frames.ctypes.data_as(C.c_void_p), len(frames),
                                         C.byref(nf), stats.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError("orc_link_decode: frame buffer overflow")
    return frames[:nf.value], stats, rows


def link_time(kind: str, blocks: np.ndarray, reps: int) -> float:
    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DT)
    return float(_link_lib(kind).orc_link_time(blocks.ctypes.data_as(C.c_void_p), len(blocks), reps))


def time_cu8(kind: str, iq: np.ndarray, reps: int, Fr=136_975_000, Fo=-50_000, fs=2_000_000, sdrclk=500) -> float:
    """Seconds for `reps` passes of one private channel over `iq` (taps off); GIL released."""
    lib = load(kind)
    iq = np.ascontiguousarray(iq)
    return float(lib.orc_time_cu8(Fr, Fo, fs, sdrclk, iq.ctypes.data_as(C.c_void_p), iq.size // 2, reps))


# ---- row f4 (oracle only so far): the fields out() / outacars() derive from a frame before formatting (orc_avlc_api.h)
AVLC_DT = np.dtype([("faddr", "<u4"), ("taddr", "<u4"), ("fromair", "u1"), ("rep", "u1"), ("gnd", "u1"), ("lc", "u1"), ("kind", "u1"),
                    ("mode", "u1"), ("ack", "u1"), ("bid", "u1"), ("bs", "u1"), ("be", "u1"), ("label", "u1", (2,)), ("reg", "u1", (7,)),
                    ("nno", "u1"), ("nfid", "u1"), ("no", "u1", (4,)), ("fid", "u1", (6,)), ("pad", "u1"),
                    ("txt_off", "<u2"), ("txt_len", "<u2"), ("info_off", "<u2"), ("info_len", "<u2")])
assert AVLC_DT.itemsize == 48
AVLC_KINDS = ("empty", "xid", "acars", "acars_badcrc", "other")
_AVLC = {}


def _avlc_lib(kind: str):
    if kind not in _AVLC:
        path = os.path.join(HERE, "libvdl2avlcport.so") if kind == "port" else os.path.join(HERE, "_ref", "libvdl2outref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"oracle library {path} missing (run `make -C oracle`)")
        lib = C.CDLL(path)
        if kind == "port":
            lib.orc_avlc_extract.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        else:
            lib.orc_out_json.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_char_p, C.c_int]
            lib.orc_out_text.argtypes = lib.orc_out_json.argtypes
        _AVLC[kind] = lib
    return _AVLC[kind]


def avlc_extract(hdata: bytes) -> np.ndarray:
    """Port: the binary record of one frame (hdata including both flags, as handed to out(); vdlm2.h:134)."""
    rec = np.zeros(1, AVLC_DT)
    buf = (C.c_uint8 * (len(hdata) + 64)).from_buffer_copy(bytes(hdata) + bytes(64))
    _avlc_lib("port").orc_avlc_extract(buf, len(hdata), rec.ctypes.data_as(C.c_void_p))
    return rec[0]


def out_json(hdata: bytes, chn: int = 0, Fr: int = 136_975_000, ppm: float = 0.0, t: float = 0.0) -> str:
    """Reference: the JSON line its out() prints for this frame with -J -G -E ('' when it prints none)."""
    buf = C.create_string_buffer(60000)
    src = (C.c_uint8 * len(hdata)).from_buffer_copy(bytes(hdata))
    n = _avlc_lib("ref").orc_out_json(src, len(hdata), chn, Fr, C.c_float(ppm), C.c_double(t), buf, len(buf))
    return buf.raw[:n].decode("latin-1")


def out_text(hdata: bytes, chn: int = 0, Fr: int = 136_975_000, ppm: float = 0.0, t: float = 0.0) -> str:
    """Reference: the text its out() prints for this frame at the default verbosity with -G -E -U."""
    buf = C.create_string_buffer(60000)
    src = (C.c_uint8 * len(hdata)).from_buffer_copy(bytes(hdata))
    n = _avlc_lib("ref").orc_out_text(src, len(hdata), chn, Fr, C.c_float(ppm), C.c_double(t), buf, len(buf))
    return buf.raw[:n].decode("latin-1")


def out_time(kind: str, data: np.ndarray, off: np.ndarray, length: np.ndarray, reps: int) -> float:
    """Seconds for `reps` passes over packed frames through the reference's out() in JSON mode (kind "ref", one core), or through
    the port's field walk (kind "port")."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    length = np.ascontiguousarray(length, dtype=np.int32)
    if kind == "ref":
        lib = _avlc_lib("ref")
        lib.orc_out_time.restype = C.c_double
        lib.orc_out_time.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        return float(lib.orc_out_time(data.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), length.ctypes.data_as(C.c_void_p),
                                      len(off), reps))
    import time
    lib = _avlc_lib("port")
    rec = np.zeros(1, AVLC_DT)
    t0 = time.perf_counter()
    for _ in range(reps):
        for o, l in zip(off.tolist(), length.tolist()):
            lib.orc_avlc_extract(data[o:o + l + 64].ctypes.data_as(C.c_void_p), l, rec.ctypes.data_as(C.c_void_p))
    return time.perf_counter() - t0


def out_available(kind: str) -> bool:
    return os.path.exists(os.path.join(HERE, "libvdl2avlcport.so") if kind == "port" else os.path.join(HERE, "_ref", "libvdl2outref.so"))
