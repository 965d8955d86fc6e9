"""Host build + ctypes front end of the warp emulator (TEST INFRASTRUCTURE, see vdl2_emul.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.pyoracle import BLOCK_DT, STEP_DT, SYM_DT, SYNC_DT

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libvdl2emul.so")
_lib = None


AVLC_LIB = os.path.join(HERE, "libvdl2avlcemul.so")
_avlc = None


def build_avlc():
    src = os.path.join(HERE, "avlc_host.cpp")
    hdr = os.path.join(ROOT, "vdlm2dec_b200", "csrc", "vdl2_avlc.cuh")
    if os.path.exists(AVLC_LIB) and all(os.path.getmtime(AVLC_LIB) >= os.path.getmtime(d) for d in (src, hdr)):
        return AVLC_LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-shared", f"-I{os.path.join(ROOT, 'vdlm2dec_b200', 'csrc')}",
                    "-o", AVLC_LIB, src], check=True)
    return AVLC_LIB


def avlc(frames: np.ndarray) -> np.ndarray:
    """The row-f4 kernel's per-frame walk (vdl2_avlc.cuh) on the host: frames (FRAME_DT) -> 48-byte records."""
    global _avlc
    from oracle.pyoracle import AVLC_DT, FRAME_DT
    if _avlc is None:
        _avlc = C.CDLL(build_avlc())
    frames = np.ascontiguousarray(frames, dtype=FRAME_DT)
    recs = np.zeros(len(frames), AVLC_DT)
    _avlc.emul_avlc(frames.ctypes.data_as(C.c_void_p), len(frames), recs.ctypes.data_as(C.c_void_p))
    return recs


MMA_LIB = os.path.join(HERE, "libvdl2mmaemul.so")
_mma = None


def build_mma():
    src = os.path.join(HERE, "mma_mix_host.cpp")
    csrc = os.path.join(ROOT, "vdlm2dec_b200", "csrc")
    deps = [src, os.path.join(csrc, "vdl2_mma_tables.h"), os.path.join(csrc, "vdl2_common.h")]
    if os.path.exists(MMA_LIB) and all(os.path.getmtime(MMA_LIB) >= os.path.getmtime(d) for d in deps):
        return MMA_LIB
    cuda_inc = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-ffp-contract=off", "-fPIC", "-shared", f"-I{cuda_inc}", f"-I{csrc}",
                    "-o", MMA_LIB, src], check=True)
    return MMA_LIB


def mma_mix(rows: np.ndarray, Fo: int, cu8: bool = True, fs: int = 2_000_000, sdrclk: int = 500):
    """Lane-level replay of the int8 tensor-core mixer (mix_rows_mma) on the product's host tables: rows = (<= 32, 2 * fs/1000)
    8-bit IQ bytes -> (complex64 dumps [32 * 84] in time order, protocol error count)."""
    global _mma
    if _mma is None:
        _mma = C.CDLL(build_mma())
    rows = np.ascontiguousarray(rows)
    assert rows.ndim == 2 and rows.shape[0] <= 32 and rows.shape[1] == 2 * (fs // 1000) and rows.dtype.itemsize == 1
    out = np.zeros((32, 84, 2), np.float32)
    rc = _mma.emul_mma_mix(rows.ctypes.data_as(C.c_void_p), rows.shape[0], C.c_uint(fs), C.c_uint(sdrclk), int(Fo), int(cu8),
                           out.ctypes.data_as(C.c_void_p))
    return (out[..., 0] + 1j * out[..., 1]).astype(np.complex64).reshape(-1), rc


def nco_table(Fo: int, fs: int) -> np.ndarray:
    """The product's oscillator table for (Fo, fs): vdl2_nco_table of vdl2_mma_tables.h, what vdl2_host.cu uploads."""
    global _mma
    if _mma is None:
        _mma = C.CDLL(build_mma())
    out = np.zeros(2 * (fs // 25000), np.float32)
    _mma.emul_nco_table(int(Fo), C.c_uint(fs), out.ctypes.data_as(C.c_void_p))
    return out.view(np.complex64)


def build():
    build_avlc()
    build_mma()
    src = os.path.join(HERE, "emul_main.cpp")
    deps = [src, os.path.join(HERE, "vdl2_emul.h")] + [os.path.join(ROOT, "vdlm2dec_b200", "csrc", f)
                                                        for f in ("vdl2_demod.cuh", "vdl2_common.h", "vdl2_tables.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cuda_inc = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", f"-I{cuda_inc}",
                    f"-I{os.path.join(ROOT, 'vdlm2dec_b200', 'csrc')}", f"-I{HERE}", "-o", LIB, src], check=True)
    return LIB


def burst_pre_symbols(wrong: bool = False) -> int:
    """Symbols whose phase the emulated warp computed ahead of the chain so far (BurstPre), on the true grid / on a wrong one."""
    f = _lib.emul_burst_pre_symbols
    f.restype = C.c_long
    return int(f(1 if wrong else 0))


def demod(dumps: np.ndarray, tile_dumps: int = 2688, chn: int = 0, Fr: int = 136_975_000, flags: int = 0, want_steps: bool = True):
    """Run the kernel's phase 2 source on the host over a decimated stream; returns (blocks, steps, syncs, syms)."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    d = np.ascontiguousarray(dumps.astype(np.complex64)).view(np.float32)
    n = len(dumps)
    blocks = np.zeros(256, BLOCK_DT)
    steps = np.zeros(n // 2 + 64, STEP_DT)
    syncs = np.zeros(1024, SYNC_DT)
    syms = np.zeros(n // 8 + 64, SYM_DT)
    nb, nst, nsy, nsm = C.c_uint(0), C.c_uint(0), C.c_uint(0), C.c_uint(0)
    rc = _lib.emul_demod(d.ctypes.data_as(C.c_void_p), C.c_long(n), int(tile_dumps), chn, Fr, C.c_uint(flags),
                         blocks.ctypes.data_as(C.c_void_p), len(blocks), C.byref(nb),
                         steps.ctypes.data_as(C.c_void_p) if want_steps else None, len(steps), C.byref(nst),
                         syncs.ctypes.data_as(C.c_void_p), len(syncs), C.byref(nsy),
                         syms.ctypes.data_as(C.c_void_p), len(syms), C.byref(nsm))
    if rc:
        raise RuntimeError(f"emul_demod failed ({rc})")
    return blocks[:nb.value], steps[:nst.value], syncs[:nsy.value], syms[:nsm.value]
