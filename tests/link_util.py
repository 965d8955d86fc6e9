"""Shared generator of completed blocks for the link-layer tests (block pipeline of vdlm2.c:84-161)."""
import numpy as np

from oracle.pyoracle import BLOCK_DT
from vdlm2dec_b200 import synth


def block_from_burst(b: synth.Burst, chn: int = 0, sync_dump: int = 0) -> np.ndarray:
    blk = np.zeros((), BLOCK_DT)
    blk["chn"], blk["Fr"], blk["ppm"] = chn, 136_975_000, 1.5
    blk["nbrow"], blk["nlbyte"] = b.nbrow, b.nlbyte
    blk["sync_dump"], blk["end_dump"] = sync_dump, sync_dump + 1000
    blk["data"] = b.expected_data
    return blk


def make_blocks(seed: int, n: int, kinds=("clean", "errors", "heavy", "garbage", "multi", "stuff")) -> np.ndarray:
    """n blocks cycling through `kinds`:
    clean    valid burst as transmitted            errors   1..3 byte errors per row (correctable)
    heavy    4..8 byte errors in some rows (beyond the code, the reference passes them through)
    garbage  random rows and lengths                multi    several HDLC frames in one burst
    stuff    payload rich in 0xff / 0x7d / 0x3e runs (bit stuffing, flag look-alikes)"""
    rng = np.random.default_rng(seed)
    out = np.zeros(n, BLOCK_DT)
    for i in range(n):
        kind = kinds[i % len(kinds)]
        nbytes = int(rng.choice([14, 28, 29, 35, 66, 67, 70, 200, 247, 249, 250, 251, 300, 497, 600, 1100, 1900]))
        if kind == "multi":
            bits = np.concatenate([synth.hdlc_bits(synth.random_payload(rng, int(rng.integers(12, 120)))) for _ in range(3)])
            b = synth.Burst(bits)
        elif kind == "stuff":
            p = rng.choice(np.array([0xFF, 0x7D, 0x3E, 0xF8, 0x1F, 0x00, 0xFE], np.uint8), size=nbytes)
            b = synth.Burst(synth.hdlc_bits(bytes(p)))
        else:
            b = synth.make_burst(rng, nbytes)
        if not b.valid:      # too long after bit stuffing: more than 8 rows
            b = synth.make_burst(rng, 600)
        blk = block_from_burst(b, chn=i % 8, sync_dump=1000 * i)
        d = blk["data"]
        if kind == "errors":
            for r in range(b.nbrow):
                for _ in range(int(rng.integers(1, 4))):
                    d[r, int(rng.integers(0, 255))] ^= np.uint8(rng.integers(1, 256))
        elif kind == "heavy":
            for r in range(b.nbrow):
                for _ in range(int(rng.integers(0, 9))):
                    d[r, int(rng.integers(0, 255))] ^= np.uint8(rng.integers(1, 256))
        elif kind == "garbage":
            blk["nbrow"] = int(rng.integers(1, 9))
            blk["nlbyte"] = int(rng.integers(0, 250))
            d[...] = rng.integers(0, 256, size=d.shape, dtype=np.uint8)
            if rng.random() < 0.5:      # mostly zeros with a few symbols: low-weight error patterns
                d[...] = 0
                for _ in range(int(rng.integers(1, 8))):
                    d[int(rng.integers(0, 8)), int(rng.integers(0, 255))] = np.uint8(rng.integers(1, 256))
        blk["data"] = d
        out[i] = blk
    return out
