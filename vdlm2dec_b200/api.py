"""Python mirror of the C ABI in include/vdl2gpu.h (ctypes, no torch types in the boundary).

The reference is a C program without a plugin API; its seam for this path is the object
d8psk.o (vdlm2.h:113-114).  This module is the thin host-side mirror used by the parity
tests and the bench: the names follow the reference (thread_param_t -> ChanParam with
chn/Fr/Fo, msgblk_t -> BLOCK_DT with ppm/nbrow/nlbyte/data).

There is no CPU fallback: importing works without a GPU (so the symbol checks can run on
a CPU box), but constructing a Vdl2Gpu needs an sm_100 device and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VDL2_LIB") or os.path.join(HERE, "libvdl2gpu.so")  # VDL2_LIB: experimental variants (tools/)

FORMATS = {"cu8": 0, "cs8": 1, "cs16": 2, "cf32": 3, "f32real": 4}
FMT_DTYPE = {"cu8": np.uint8, "cs8": np.int8, "cs16": np.int16, "cf32": np.float32, "f32real": np.float32}
FMT_BYTES = {"cu8": 2, "cs8": 2, "cs16": 4, "cf32": 8, "f32real": 4}
TAP_DUMPS, TAP_STEPS, TAP_SYNCS, TAP_SYMS = 1, 2, 4, 8
OPT_EXACT_IDLE = 0x100
OPT_OVERLAP = 0x400  # consecutive launches may overlap (PDL); no per-launch kernel time
OPT_DP4A_MIX = 0x800  # cu8/cs8 at 2 Msps: round-1 IDP.4A mixer instead of the int8 tensor-core mixer (A/B, parity tests)
OPT_FLOAT_MIX = 0x200  # cu8/cs8: generic fp32 mixer instead of the integer dot-product mixer (A/B, parity tests)

STEP_DT = np.dtype([("dump", "<i8"), ("P", "<f4"), ("err", "<f4"), ("fr", "<f4"), ("pad", "<i4")])
SYNC_DT = np.dtype([("dump", "<i8"), ("clk", "<i4"), ("df", "<f4"), ("ppm", "<f4"), ("P1", "<f4")])
SYM_DT = np.dtype([("dump", "<i8"), ("D", "<f4"), ("P", "<f4"), ("gi", "<i4"), ("v", "<f4", (3,)),
                   ("state_after", "<i4"), ("pad", "<i4")])
BLOCK_DT = np.dtype([("sync_dump", "<i8"), ("end_dump", "<i8"), ("chn", "<i4"), ("Fr", "<i4"), ("ppm", "<f4"),
                     ("nbrow", "<i4"), ("nlbyte", "<i4"), ("data", "u1", (8, 255)), ("pad", "u1", (4,))])
assert BLOCK_DT.itemsize == 2080
# what the reference hands to out(blk, hdata, l) (vdlm2.h:134) after rs(), HDLC un-stuffing and the FCS check
FRAME_DT = np.dtype([("block", "<i4"), ("len", "<i4"), ("chn", "<i4"), ("Fr", "<i4"), ("ppm", "<f4"), ("pad", "<i4"),
                     ("sync_dump", "<i8"), ("hdata", "u1", (2016,))])
BLKSTAT_DT = np.dtype([("rs", "i1", (8,)), ("nbytes", "<i4"), ("nframes", "<i4")])
assert FRAME_DT.itemsize == 2048 and BLKSTAT_DT.itemsize == 16
# vdl2_avlc_t (row f4): the fields out() / outacars() derive from a frame before formatting
AVLC_DT = np.dtype([("faddr", "<u4"), ("taddr", "<u4"), ("fromair", "u1"), ("rep", "u1"), ("gnd", "u1"), ("lc", "u1"), ("kind", "u1"),
                    ("mode", "u1"), ("ack", "u1"), ("bid", "u1"), ("bs", "u1"), ("be", "u1"), ("label", "u1", (2,)), ("reg", "u1", (7,)),
                    ("nno", "u1"), ("nfid", "u1"), ("no", "u1", (4,)), ("fid", "u1", (6,)), ("pad", "u1"),
                    ("txt_off", "<u2"), ("txt_len", "<u2"), ("info_off", "<u2"), ("info_len", "<u2")])
assert AVLC_DT.itemsize == 48
# vdl2_frame_hdr_t: header of a packed frame (vdl2_drain_frames_packed)
FRAME_HDR_DT = np.dtype([("sync_dump", "<i8"), ("chn", "<i4"), ("Fr", "<i4"), ("ppm", "<f4"), ("len", "<i4"), ("offset", "<u4"), ("dur", "<i4")])
assert FRAME_HDR_DT.itemsize == 32


class ChanParam(C.Structure):  # thread_param_t, vdlm2.h:49-52
    _fields_ = [("chn", C.c_int), ("Fr", C.c_int), ("Fo", C.c_int)]


class Config(C.Structure):
    _fields_ = [("fs", C.c_uint), ("sdrclk", C.c_uint), ("format", C.c_int), ("nch", C.c_int),
                ("ch_per_stream", C.c_int), ("device", C.c_int), ("taps", C.c_uint),
                ("max_samples", C.c_size_t), ("max_blocks", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("samples_in", C.c_uint64), ("samples_done", C.c_uint64),
                ("blocks_out", C.c_uint64), ("blocks_dropped", C.c_uint64), ("last_kernel_ms", C.c_float),
                ("n_sm", C.c_int), ("grid", C.c_int), ("smem_bytes", C.c_int), ("last_link_ms", C.c_float),
                ("link_launches", C.c_uint32), ("frames_out", C.c_uint64)]


EXPORTS = ["vdl2_abi_version", "vdl2_last_error", "vdl2_create", "vdl2_destroy", "vdl2_process_host",
           "vdl2_process_device", "vdl2_sync", "vdl2_drain_blocks", "vdl2_read_dumps", "vdl2_read_steps",
           "vdl2_read_syncs", "vdl2_read_syms", "vdl2_get_stats", "vdl2_cuda_stream", "vdl2_link_decode",
           "vdl2_drain_frames", "vdl2_host_alloc", "vdl2_host_free", "vdl2_process_host_rtl", "vdl2_avlc_extract",
           "vdl2_submit_host", "vdl2_submit_copy", "vdl2_pending_blocks", "vdl2_drain_frames_packed", "vdl2_last_pack_ms", "vdl2_channelise_device",
           "vdl2_multi_create", "vdl2_multi_destroy", "vdl2_multi_process_host", "vdl2_multi_drain_blocks", "vdl2_multi_ndev",
           "vdl2_multi_handle", "vdl2_multi_last_error"]

_lib = None


def load_library():
    """dlopen libvdl2gpu.so; raises (loudly) when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m vdlm2dec_b200.build` "
                           "(there is no CPU fallback for the VDL2 front end)")
    lib = C.CDLL(LIB_PATH)
    lib.vdl2_abi_version.restype = C.c_int
    lib.vdl2_last_error.restype = C.c_char_p
    lib.vdl2_last_error.argtypes = [C.c_void_p]
    lib.vdl2_create.argtypes = [C.POINTER(Config), C.POINTER(ChanParam), C.POINTER(C.c_void_p)]
    lib.vdl2_destroy.argtypes = [C.c_void_p]
    lib.vdl2_process_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.vdl2_process_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.vdl2_submit_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.vdl2_submit_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.vdl2_pending_blocks.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.vdl2_drain_frames_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_size_t,
                                             C.POINTER(C.c_size_t), C.c_void_p]
    lib.vdl2_channelise_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.vdl2_last_pack_ms.restype = C.c_float
    lib.vdl2_last_pack_ms.argtypes = [C.c_void_p]
    lib.vdl2_multi_create.argtypes = [C.POINTER(Config), C.POINTER(ChanParam), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    lib.vdl2_multi_destroy.argtypes = [C.c_void_p]
    lib.vdl2_multi_process_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.vdl2_multi_drain_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.vdl2_multi_ndev.argtypes = [C.c_void_p]
    lib.vdl2_multi_handle.restype = C.c_void_p
    lib.vdl2_multi_handle.argtypes = [C.c_void_p, C.c_int]
    lib.vdl2_multi_last_error.restype = C.c_char_p
    lib.vdl2_multi_last_error.argtypes = [C.c_void_p]
    lib.vdl2_sync.argtypes = [C.c_void_p]
    lib.vdl2_drain_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    for f in ("vdl2_read_dumps", "vdl2_read_steps", "vdl2_read_syncs", "vdl2_read_syms"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.vdl2_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.vdl2_link_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    lib.vdl2_drain_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.vdl2_avlc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.vdl2_process_host_rtl.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.vdl2_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    lib.vdl2_host_free.argtypes = [C.c_void_p]
    lib.vdl2_cuda_stream.restype = C.c_void_p
    lib.vdl2_cuda_stream.argtypes = [C.c_void_p]
    _lib = lib
    return lib


class Vdl2Error(RuntimeError):
    pass


class Vdl2Gpu:
    """A set of channels demodulated on one B200: feed IQ, drain completed blocks.

    chans: sequence of (chn, Fr, Fo) -- the reference's thread_param_t per channel.
    Channels c*ch_per_stream .. (c+1)*ch_per_stream-1 are demodulated from input stream c.
    """

    def __init__(self, chans: Sequence[tuple[int, int, int]], fs: int = 2_000_000, sdrclk: int = 500,
                 fmt: str = "cu8", ch_per_stream: int = 1, device: int = 0, taps: int = 0,
                 max_samples: int = 1 << 22, max_blocks: int = 0):
        self.lib = load_library()
        self.fmt = fmt
        self.nch = len(chans)
        self.ch_per_stream = ch_per_stream
        self.nstreams = self.nch // ch_per_stream
        self.bps = FMT_BYTES[fmt]
        self.max_samples = max_samples
        arr = (ChanParam * self.nch)(*[ChanParam(*c) for c in chans])
        cfg = Config(fs, sdrclk, FORMATS[fmt], self.nch, ch_per_stream, device, taps, max_samples, max_blocks)
        self.h = C.c_void_p()
        if self.lib.vdl2_create(C.byref(cfg), arr, C.byref(self.h)):
            raise Vdl2Error(self.lib.vdl2_last_error(None).decode())
        self._cap_blocks = max_blocks if max_blocks > 0 else max(4096, self.nch * 8)

    def _check(self, rc):
        if rc:
            raise Vdl2Error(self.lib.vdl2_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.vdl2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- input
    def process(self, iq: np.ndarray):
        """Host buffer: [nstreams, nsamples*k] (or flat for one stream), dtype of the format."""
        iq = np.ascontiguousarray(iq, dtype=FMT_DTYPE[self.fmt])
        if iq.ndim == 1:
            iq = iq.reshape(1, -1)
        assert iq.shape[0] == self.nstreams, (iq.shape, self.nstreams)
        row_bytes = iq.shape[1] * iq.itemsize
        nsamples = row_bytes // self.bps
        self._check(self.lib.vdl2_process_host(self.h, iq.ctypes.data_as(C.c_void_p), nsamples, row_bytes))
        return self

    def process_rtl(self, cu8: np.ndarray):
        """Raw cu8 bytes of whole 65536-byte RTL callbacks, demodulated as the reference's in_callback lays them out
        (rtl.c:285-292), expanded on the device; the handle must have fmt="cf32" and one stream."""
        cu8 = np.ascontiguousarray(cu8, dtype=np.uint8)
        self._check(self.lib.vdl2_process_host_rtl(self.h, cu8.ctypes.data_as(C.c_void_p), cu8.size // 2))

    def process_ptr(self, host_ptr: int, nsamples: int, pitch_bytes: int):
        self._check(self.lib.vdl2_process_host(self.h, C.c_void_p(host_ptr), nsamples, pitch_bytes))

    def submit_copy(self, iq: np.ndarray):
        """Asynchronous ingest: the samples are copied into a page-locked ring slot of the handle and the upload + launch are
        enqueued; returns at once, `iq` may be reused.  Collect with pending_blocks() / drain_blocks()."""
        iq = np.ascontiguousarray(iq, dtype=FMT_DTYPE[self.fmt])
        if iq.ndim == 1:
            iq = iq.reshape(1, -1)
        row_bytes = iq.shape[1] * iq.itemsize
        self._check(self.lib.vdl2_submit_copy(self.h, iq.ctypes.data_as(C.c_void_p), row_bytes // self.bps, row_bytes))

    def pending_blocks(self) -> int:
        n = C.c_int(0)
        self._check(self.lib.vdl2_pending_blocks(self.h, C.byref(n)))
        return n.value

    def process_device(self, dev_ptr: int, nsamples: int, pitch_bytes: int):
        """Device-resident input (e.g. a torch tensor's data_ptr()); asynchronous, see sync()."""
        self._check(self.lib.vdl2_process_device(self.h, C.c_void_p(dev_ptr), nsamples, pitch_bytes))

    def channelise_device(self, dev_ptr: int, nsamples: int, pitch_bytes: int, out_ptr: int, out_pitch: int):
        """Row f3: one pass over every stream -> decimated 84 ksps streams of all channels at out_ptr (device, complex64
        [nch, out_pitch]); asynchronous."""
        self._check(self.lib.vdl2_channelise_device(self.h, C.c_void_p(dev_ptr), nsamples, pitch_bytes, C.c_void_p(out_ptr), out_pitch))

    def sync(self):
        self._check(self.lib.vdl2_sync(self.h))

    # ---- output
    def drain_blocks(self) -> np.ndarray:
        out = np.zeros(self._cap_blocks, dtype=BLOCK_DT)
        n = C.c_int(0)
        self._check(self.lib.vdl2_drain_blocks(self.h, out.ctypes.data_as(C.c_void_p), len(out), C.byref(n)))
        return out[:n.value].copy()

    def link_decode(self, blocks: np.ndarray, want_rows: bool = True):
        """Block pipeline behind the demodulator (blk_thread, vdlm2.c:84-161) on the GPU for blocks in host memory:
        returns (frames, per-block stats, data rows after rs())."""
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DT)
        n = len(blocks)
        frames = np.zeros(max(4 * n, 16), FRAME_DT)
        stats = np.zeros(n, BLKSTAT_DT)
        rows = np.zeros((n, 8, 255), np.uint8) if want_rows else None
        nf = C.c_int(0)
        self._check(self.lib.vdl2_link_decode(self.h, blocks.ctypes.data_as(C.c_void_p), n, frames.ctypes.data_as(C.c_void_p),
                                              len(frames), C.byref(nf), stats.ctypes.data_as(C.c_void_p),
                                              rows.ctypes.data_as(C.c_void_p) if want_rows else None))
        return frames[:nf.value].copy(), stats, rows

    def avlc_extract(self, frames: np.ndarray) -> np.ndarray:
        """Row f4: one field record per frame (addresses, direction, payload class, ACARS header fields and text extent)."""
        frames = np.ascontiguousarray(frames, dtype=FRAME_DT)
        recs = np.zeros(len(frames), AVLC_DT)
        self._check(self.lib.vdl2_avlc_extract(self.h, frames.ctypes.data_as(C.c_void_p), len(frames), recs.ctypes.data_as(C.c_void_p)))
        return recs

    def drain_frames(self):
        """Completed blocks -> block pipeline on the device -> (frames, blocks); frame['block'] indexes blocks."""
        blocks = np.zeros(self._cap_blocks, dtype=BLOCK_DT)
        frames = np.zeros(max(2 * self._cap_blocks, 16), FRAME_DT)
        nf, nb = C.c_int(0), C.c_int(0)
        self._check(self.lib.vdl2_drain_frames(self.h, frames.ctypes.data_as(C.c_void_p), len(frames), C.byref(nf),
                                               blocks.ctypes.data_as(C.c_void_p), len(blocks), C.byref(nb)))
        return frames[:nf.value].copy(), blocks[:nb.value].copy()

    def drain_frames_packed(self, want_records: bool = True, max_frames: int | None = None):
        """Rows f1 + f4 end to end: (headers, packed frame bytes, field records or None), completion order; page-locked buffers
        owned by the handle (allocated once with vdl2_host_alloc)."""
        nf_cap = max_frames or max(2 * self._cap_blocks, 16)
        if getattr(self, "_pk", None) is None or self._pk[0] < nf_cap:
            def pinned(nbytes, dt):
                p = C.c_void_p()
                if self.lib.vdl2_host_alloc(nbytes, C.byref(p)):
                    raise Vdl2Error(self.lib.vdl2_last_error(None).decode())
                buf = (C.c_char * nbytes).from_address(p.value)
                return np.frombuffer(buf, dtype=dt)
            self._pk = (nf_cap, pinned(32 * nf_cap, FRAME_HDR_DT), pinned(512 * nf_cap, np.uint8), pinned(48 * nf_cap, AVLC_DT))
        _, hdrs, data, recs = self._pk
        nf, nb = C.c_int(0), C.c_size_t(0)
        self._check(self.lib.vdl2_drain_frames_packed(self.h, hdrs.ctypes.data_as(C.c_void_p), nf_cap, C.byref(nf), data.ctypes.data_as(C.c_void_p),
                                                      data.nbytes, C.byref(nb), recs.ctypes.data_as(C.c_void_p) if want_records else None))
        return hdrs[:nf.value], data[:nb.value], (recs[:nf.value] if want_records else None)

    @property
    def last_pack_ms(self) -> float:
        return float(self.lib.vdl2_last_pack_ms(self.h))

    def _read(self, fn, ch, dt, cap):
        out = np.zeros(cap, dtype=dt)
        n = C.c_size_t(0)
        self._check(fn(self.h, ch, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return out[:n.value].copy()

    def _rows_cap(self):
        return self.max_samples // 1000 * 84 // (2_000_000 // 1000) + 4096

    def read_dumps(self, ch: int) -> np.ndarray:
        cap = (self.max_samples // 20 + 4096)
        return self._read(self.lib.vdl2_read_dumps, ch, np.dtype("<c8"), cap)

    def read_steps(self, ch: int) -> np.ndarray:
        return self._read(self.lib.vdl2_read_steps, ch, STEP_DT, self.max_samples // 40 + 4096)

    def read_syncs(self, ch: int) -> np.ndarray:
        return self._read(self.lib.vdl2_read_syncs, ch, SYNC_DT, self.max_samples // 1000 + 4096)

    def read_syms(self, ch: int) -> np.ndarray:
        return self._read(self.lib.vdl2_read_syms, ch, SYM_DT, self.max_samples // 160 + 4096)

    def stats(self) -> dict:
        st = Stats()
        self._check(self.lib.vdl2_get_stats(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in Stats._fields_}

    @property
    def cuda_stream(self) -> int:
        return int(self.lib.vdl2_cuda_stream(self.h) or 0)


class Vdl2Multi:
    """Several GPUs behind one handle (vdl2_multi_*, include/vdl2gpu.h): stream s with its channels lives on devices[s mod N],
    no collective; completed blocks come back merged in the order one GPU would have produced them."""

    def __init__(self, chans: Sequence[tuple[int, int, int]], devices: Sequence[int], fs: int = 2_000_000, sdrclk: int = 500,
                 fmt: str = "cu8", ch_per_stream: int = 1, taps: int = 0, max_samples: int = 1 << 22, max_blocks: int = 0):
        self.lib = load_library()
        self.fmt = fmt
        self.nch = len(chans)
        self.nstreams = self.nch // ch_per_stream
        self.bps = FMT_BYTES[fmt]
        arr = (ChanParam * self.nch)(*[ChanParam(*c) for c in chans])
        dev = (C.c_int * len(devices))(*devices)
        cfg = Config(fs, sdrclk, FORMATS[fmt], self.nch, ch_per_stream, 0, taps, max_samples, max_blocks)
        self.h = C.c_void_p()
        if self.lib.vdl2_multi_create(C.byref(cfg), arr, dev, len(devices), C.byref(self.h)):
            raise Vdl2Error(self.lib.vdl2_multi_last_error(None).decode())
        self._cap_blocks = (max_blocks if max_blocks > 0 else max(4096, self.nch * 8)) * len(devices)

    def _check(self, rc):
        if rc:
            raise Vdl2Error(self.lib.vdl2_multi_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.vdl2_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ndev(self) -> int:
        return self.lib.vdl2_multi_ndev(self.h)

    def process(self, iq: np.ndarray):
        iq = np.ascontiguousarray(iq, dtype=FMT_DTYPE[self.fmt])
        if iq.ndim == 1:
            iq = iq.reshape(1, -1)
        assert iq.shape[0] == self.nstreams, (iq.shape, self.nstreams)
        row_bytes = iq.shape[1] * iq.itemsize
        self._check(self.lib.vdl2_multi_process_host(self.h, iq.ctypes.data_as(C.c_void_p), row_bytes // self.bps, row_bytes))
        return self

    def process_ptr(self, host_ptr: int, nsamples: int, pitch_bytes: int):
        self._check(self.lib.vdl2_multi_process_host(self.h, C.c_void_p(host_ptr), nsamples, pitch_bytes))

    def drain_blocks(self) -> np.ndarray:
        out = np.zeros(self._cap_blocks, dtype=BLOCK_DT)
        n = C.c_int(0)
        self._check(self.lib.vdl2_multi_drain_blocks(self.h, out.ctypes.data_as(C.c_void_p), len(out), C.byref(n)))
        return out[:n.value].copy()
