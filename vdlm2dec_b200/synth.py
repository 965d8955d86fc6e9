"""Synthetic VDL Mode 2 transmitter (numpy) -- seeded test/bench input, not part of the hot path.

The reference ships no recorded IQ and no golden vectors, so every parity test and the
bench feed on bursts built here.  The transmit chain is the inverse of what the reference
receiver accepts; each step cites the receiver code it inverts:

  AVLC frame  : 0x7e | bit-stuffed(payload | PPP-FCS16 LE) | 0x7e, LSB first   (vdlm2.c:120-153, :51-55)
  rows        : 1992-bit rows -> 249 bytes + 6 RS(255,249) parity bytes, GF(256)
                poly 0x187, generator roots alpha^120..125                      (rs.c:17-19,75-76)
  header      : 0,0,0, len[0..16] LSB first, 5 parity bits                      (d8psk.c:81-95, viterbi.c:29-35)
  interleave  : column-major over rows, short last row, 0/2/4/6 FEC bytes there (d8psk.c:139-162,188-197)
  scrambler   : additive, seed 0x4D4B, taps 0 and 14                            (d8psk.c:54-65,299)
  symbols     : 3 bits -> Gray -> differential pi/4 steps                       (d8psk.h Grey tables)
  preamble    : reference symbol + 16-symbol unique word                        (d8psk.h:20-26)
  pulse       : raised cosine alpha 0.6, 10500 sym/s, mixed to +Fo, quantised   (d8psk.h:28-45, rtl.c:287-289)
"""
from __future__ import annotations

import numpy as np

SYMRATE = 10500
ROWBITS = 1992
# differential indices (units of pi/4) of the 16-symbol unique word, SW[l] - l*pi/8 (d8psk.h:20-26)
UNIQUE_WORD = [0, 3, 2, 4, 0, 1, 6, 4, 1, 7, 2, 5, 6, 5, 7, 3]
# 3 bits (first = MSB) -> differential index; inverse of the soft demap sectors (ggrey.c:60-103)
GRAY = {0b000: 0, 0b001: 1, 0b011: 2, 0b010: 3, 0b110: 4, 0b111: 5, 0b101: 6, 0b100: 7}
HCOL = [0x06, 0x07, 0x09, 0x0A, 0x0B, 0x0C, 0x0E, 0x0F, 0x11, 0x13,
        0x15, 0x16, 0x18, 0x19, 0x1A, 0x1B, 0x1C, 0x1D, 0x1E, 0x1F]


# ----------------------------------------------------------------------------- link layer
def fcs16(data: bytes) -> int:
    """PPP FCS-16 (reflected 0x8408, init 0xffff, final xor 0xffff); vdlm2.c:50-55 checks 0xf0b8."""
    crc = 0xFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
    return crc ^ 0xFFFF


def hdlc_bits(payload: bytes) -> np.ndarray:
    """Flag + bit-stuffed(payload + FCS) + flag as a bit array, LSB first (inverse of vdlm2.c:120-153)."""
    f = fcs16(payload)
    body = payload + bytes([f & 0xFF, f >> 8])
    bits = [(0x7E >> i) & 1 for i in range(8)]
    ones = 0
    for byte in body:
        for i in range(8):
            b = (byte >> i) & 1
            bits.append(b)
            if b:
                ones += 1
                if ones == 5:
                    bits.append(0)
                    ones = 0
            else:
                ones = 0
    bits += [(0x7E >> i) & 1 for i in range(8)]
    return np.array(bits, dtype=np.uint8)


def _gf_tables():
    exp = np.zeros(512, dtype=np.int32)
    log = np.zeros(256, dtype=np.int32)
    x = 1
    for i in range(255):
        exp[i] = x
        log[x] = i
        x <<= 1
        if x & 0x100:
            x ^= 0x187  # rs.c:17-19 (alpha^8 = 0x87)
    exp[255:510] = exp[0:255]
    return exp, log


_GF_EXP, _GF_LOG = _gf_tables()


def _gf_mul(a: int, b: int) -> int:
    if a == 0 or b == 0:
        return 0
    return int(_GF_EXP[_GF_LOG[a] + _GF_LOG[b]])


def _rs_genpoly():
    g = [1]
    for r in range(120, 126):  # rs.c:75-76: first consecutive root 120, 6 roots
        root = int(_GF_EXP[r])
        ng = [0] * (len(g) + 1)
        for i, c in enumerate(g):
            ng[i] ^= _gf_mul(c, root)
            ng[i + 1] ^= c
        g = ng
    return g  # low order first, monic degree 6


_RS_GEN = _rs_genpoly()


def rs_parity(row: bytes) -> bytes:
    """Systematic RS(255,249) parity; row[0] is the highest-degree coefficient (rs.c:81-291 decodes it)."""
    assert len(row) == 249
    rem = [0] * 6  # rem[5] highest
    for d in row:
        fb = d ^ rem[5]
        for j in range(5, 0, -1):
            rem[j] = rem[j - 1] ^ _gf_mul(fb, _RS_GEN[j])
        rem[0] = _gf_mul(fb, _RS_GEN[0])
    return bytes(rem[::-1])


# ----------------------------------------------------------------------------- burst bits
def header_bits(length: int) -> np.ndarray:
    """25 header bits: 0,0,0, len LSB first (17 bits), 5 parity bits MSB first (viterbi.c:29-35,80-96)."""
    bits = [0, 0, 0] + [(length >> i) & 1 for i in range(17)]
    syn = 0
    for n, b in enumerate(bits):
        if b:
            syn ^= HCOL[n]
    bits += [(syn >> (4 - i)) & 1 for i in range(5)]
    return np.array(bits, dtype=np.uint8)


def fec_bytes_last_row(nlbyte: int) -> int:
    """How many parity bytes the short last row carries (d8psk.c:153-161 / vdlm2.c:64-82)."""
    if nlbyte <= 2:
        return 0
    if nlbyte <= 30:
        return 2
    if nlbyte <= 67:
        return 4
    return 6


class Burst:
    """Everything known about one transmitted burst (expected receiver outputs included)."""

    def __init__(self, frame_bits: np.ndarray, length_override: int | None = None, corrupt_rows: bool = False):
        length = int(len(frame_bits)) if length_override is None else int(length_override)
        self.length = length
        self.nbrow = length // ROWBITS + 1
        self.nlbyte = (length % ROWBITS + 7) // 8
        nrows = min(self.nbrow, 8) if self.nbrow <= 8 else self.nbrow
        padded = np.zeros(max(nrows, 1) * ROWBITS, dtype=np.uint8)
        m = min(len(frame_bits), len(padded))
        padded[:m] = frame_bits[:m]
        rows = []
        for r in range(nrows):
            rb = padded[r * ROWBITS:(r + 1) * ROWBITS].reshape(249, 8)
            data = bytes(int(sum(int(rb[i, k]) << k for k in range(8))) for i in range(249))
            rows.append(bytearray(data + rs_parity(data)))
        self.rows = rows  # full 255-byte rows (what a receiver with all bytes would hold)
        self.valid = length >= 96 and self.nbrow <= 8
        self.tx_bits = self._tx_bits()
        self.expected_data = self._expected_block() if self.valid else None

    def _byte_order(self):
        """(row, col) sequence of transmitted bytes: data phase then FEC phase (d8psk.c:117-206)."""
        nbrow, nlbyte = self.nbrow, self.nlbyte
        order = []
        for c in range(249):
            for r in range(nbrow):
                if nlbyte and r == nbrow - 1 and c >= nlbyte:
                    continue
                order.append((r, c))
        # FEC phase: the receiver rewrites nbrow/nlbyte (d8psk.c:153-161)
        if nlbyte <= 2:
            f_rows, f_last = nbrow - 1, 0
        else:
            f_rows, f_last = nbrow, fec_bytes_last_row(nlbyte)
            if f_last == 6:
                f_last = 0  # "else nlbyte = 0": full FEC on every row
        for c in range(6):
            for r in range(f_rows):
                if f_last and r == f_rows - 1 and c >= f_last:
                    continue
                order.append((r, 249 + c))
        return order

    def _tx_bits(self) -> np.ndarray:
        bits = list(header_bits(self.length & 0x1FFFF))
        if self.valid:
            for (r, c) in self._byte_order():
                v = self.rows[r][c]
                bits += [(v >> i) & 1 for i in range(8)]
        return np.array(bits, dtype=np.uint8)

    def _expected_block(self) -> np.ndarray:
        """msgblk_t.data[0..7][0..254] as the receiver leaves it (untransmitted cells stay 0)."""
        blk = np.zeros((8, 255), dtype=np.uint8)
        for (r, c) in self._byte_order():
            blk[r, c] = self.rows[r][c]
        return blk


def scramble(bits: np.ndarray) -> np.ndarray:
    """Additive scrambler over all bits from the first header bit (d8psk.c:54-65, seed :299)."""
    s = 0x4D4B
    out = np.empty_like(bits)
    for i, b in enumerate(bits):
        k = (s ^ (s >> 14)) & 1
        s = ((s << 1) | k) & 0xFFFFFFFF
        out[i] = b ^ k
    return out


def scrambler_sequence(n: int) -> np.ndarray:
    return scramble(np.zeros(n, dtype=np.uint8))


def symbols_from_bits(bits: np.ndarray) -> np.ndarray:
    """Scrambled bits -> differential indices (pi/4 units); a trailing partial symbol is zero padded."""
    pad = (-len(bits)) % 3
    b = np.concatenate([bits, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 3)
    tri = (b[:, 0] << 2) | (b[:, 1] << 1) | b[:, 2]
    lut = np.zeros(8, dtype=np.int64)
    for k, v in GRAY.items():
        lut[k] = v
    return lut[tri]


def burst_phase_indices(burst: Burst, n_ramp: int = 4, tail: int = 2, rng=None) -> np.ndarray:
    """Absolute symbol phases (pi/4 units) of ramp + reference + unique word + header/data + tail."""
    rng = rng or np.random.default_rng(0)
    diff = list(rng.integers(0, 8, size=n_ramp)) + [0] + UNIQUE_WORD
    diff += list(symbols_from_bits(scramble(burst.tx_bits)))
    diff += list(rng.integers(0, 8, size=tail))
    return np.cumsum(np.array(diff, dtype=np.int64)) % 8


# ----------------------------------------------------------------------------- waveform
def raised_cosine(x: np.ndarray, alpha: float = 0.6) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    den = 1.0 - (2.0 * alpha * x) ** 2
    sing = np.abs(den) < 1e-9
    den = np.where(sing, 1.0, den)
    y = np.sinc(x) * np.cos(np.pi * alpha * x) / den
    return np.where(sing, (np.pi / 4.0) * np.sinc(1.0 / (2.0 * alpha)), y)


def burst_waveform(phase_idx: np.ndarray, fs: float, t0: float, n0: int, n1: int, span: int = 5) -> np.ndarray:
    """Baseband samples n0..n1-1 of a burst whose symbol k peaks at sample t0 + k*fs/10500."""
    T = fs / SYMRATE
    n = np.arange(n0, n1, dtype=np.float64)
    u = (n - t0) / T
    kc = np.floor(u).astype(np.int64)
    ph = np.exp(1j * (np.pi / 4.0) * phase_idx.astype(np.float64))
    out = np.zeros(len(n), dtype=np.complex128)
    K = len(phase_idx)
    for j in range(-span + 1, span + 1):
        k = kc + j
        ok = (k >= 0) & (k < K)
        kk = np.clip(k, 0, K - 1)
        out += np.where(ok, ph[kk] * raised_cosine(u - k), 0.0)
    return out


class ChannelSpec:
    """One synthetic channel: where its bursts sit and how loud/offset they are."""

    def __init__(self, Fo: int, bursts=(), noise_sigma: float = 8.0, seed: int = 0):
        self.Fo = Fo
        self.bursts = list(bursts)  # dicts: burst, start(float sample), amp, cfo(Hz), phase0
        self.noise_sigma = noise_sigma
        self.seed = seed


def render_channel(spec: ChannelSpec, nsamples: int, fs: int = 2_000_000, fmt: str = "cu8") -> np.ndarray:
    """Complex baseband -> +Fo -> quantised samples.  fmt: cu8 | cs8 | cs16 | cf32 (interleaved IQ)."""
    rng = np.random.default_rng(spec.seed)
    x = np.zeros(nsamples, dtype=np.complex128)
    T = fs / SYMRATE
    for b in spec.bursts:
        pidx = b["phase_idx"]
        t0 = float(b["start"])
        n0 = max(0, int(np.floor(t0 - 6 * T)))
        n1 = min(nsamples, int(np.ceil(t0 + (len(pidx) + 6) * T)))
        if n1 <= n0:
            continue
        w = burst_waveform(pidx, fs, t0, n0, n1)
        n = np.arange(n0, n1, dtype=np.float64)
        rot = np.exp(1j * (2.0 * np.pi * (spec.Fo + b.get("cfo", 0.0)) * n / fs + b.get("phase0", 0.0)))
        x[n0:n1] += b.get("amp", 60.0) * w * rot
    if spec.noise_sigma > 0:
        x += spec.noise_sigma * (rng.standard_normal(nsamples) + 1j * rng.standard_normal(nsamples))
    return quantise(x, fmt)


def quantise(x: np.ndarray, fmt: str) -> np.ndarray:
    iq = np.empty(2 * len(x), dtype=np.float64)
    iq[0::2] = x.real
    iq[1::2] = x.imag
    if fmt == "cu8":
        return np.clip(np.rint(iq + 127.37), 0, 255).astype(np.uint8)
    if fmt == "cs8":
        return np.clip(np.rint(iq), -128, 127).astype(np.int8)
    if fmt == "cs16":
        return np.clip(np.rint(iq * 128.0), -32768, 32767).astype(np.int16)
    if fmt == "cf32":
        return iq.astype(np.float32)
    if fmt == "f32real":  # Airspy AIRSPY_SAMPLE_FLOAT32_REAL (air.c:123): real samples, full scale ~1
        return (x.real / 64.0).astype(np.float32)
    raise ValueError(fmt)


def random_payload(rng, nbytes: int) -> bytes:
    """Payload without 0x7e bytes (they would read as flags downstream, vdlm2.c:136)."""
    p = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
    p[p == 0x7E] = 0x7D
    return bytes(p)


def make_burst(rng, payload_bytes: int, **kw) -> Burst:
    return Burst(hdlc_bits(random_payload(rng, payload_bytes)), **kw)


def standard_channel(seed: int, nsamples: int, Fo: int, fs: int = 2_000_000, period: int | None = None,
                     payload_bytes=(30, 600), amp=(25.0, 70.0), cfo=500.0, noise_sigma: float = 8.0,
                     first: float | None = None) -> ChannelSpec:
    """The bench/test channel shape of SURVEY.md section 8(d): AWGN + periodic valid bursts with random CFO."""
    rng = np.random.default_rng(seed)
    T = fs / SYMRATE
    bursts = []
    t = float(first) if first is not None else float(rng.uniform(3000, 20000))
    while True:
        nb = int(rng.integers(payload_bytes[0], payload_bytes[1] + 1))
        b = make_burst(rng, nb)
        pidx = burst_phase_indices(b, rng=rng)
        dur = (len(pidx) + 8) * T
        if t + dur >= nsamples:
            break
        bursts.append(dict(burst=b, phase_idx=pidx, start=t + 4 * T + float(rng.uniform(0, T)),
                           amp=float(rng.uniform(*amp)), cfo=float(rng.uniform(-cfo, cfo)),
                           phase0=float(rng.uniform(0, 2 * np.pi))))
        gap = float(rng.uniform(0.2, 1.0)) * (period if period else 40000)
        t += dur + gap
    return ChannelSpec(Fo, bursts, noise_sigma=noise_sigma, seed=seed + 7919)


# ----------------------------------------------------------------------------- ACARS over AVLC
def _rev(v: int, n: int) -> int:
    return int(f"{v:0{n}b}"[::-1], 2)


def avlc_addr(addr27: int, bit1: int, last: bool) -> bytes:
    """Inverse of icaoaddr() (out.c:426-435): 27-bit address, 7 bits per byte, bit-reversed, LSB = HDLC extension."""
    b0 = (_rev((addr27 >> 21) & 0x3F, 6) << 2) | (bit1 << 1)
    b1 = _rev((addr27 >> 14) & 0x7F, 7) << 1
    b2 = _rev((addr27 >> 7) & 0x7F, 7) << 1
    b3 = (_rev(addr27 & 0x7F, 7) << 1) | (1 if last else 0)
    return bytes([b0, b1, b2, b3])


def _crc_kermit(data: bytes) -> int:
    crc = 0
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
    return crc


def acars_frame(icao: int, reg: str, label: str, text: str, msgno: str = "M01A", flight: str = "AB1234",
                ground: int = 0x10A0B0) -> bytes:
    """AVLC information frame carrying one ACARS message (what outacars.c:214-331 parses):
    dst, src (aircraft: type 1), control, ff ff 01, ACARS body with odd parity + CRC-16 + DEL.
    Returns the payload for hdlc_bits() (FCS and flags are added there)."""
    body = "2" + reg.rjust(7, ".")[:7] + "\x15" + label[:2] + "1" + "\x02" + msgno[:4] + flight[:6] + text + "\x03"
    raw = bytearray()
    for ch in body.encode("ascii"):
        ch &= 0x7F
        raw.append(ch | (0x80 if bin(ch).count("1") % 2 == 0 else 0))  # odd parity in bit 7
    crc = _crc_kermit(bytes(raw))
    raw += bytes([crc & 0xFF, crc >> 8, 0x7F])
    dst = avlc_addr((2 << 24) | ground, 0, False)
    src = avlc_addr((1 << 24) | (icao & 0xFFFFFF), 0, True)
    return dst + src + bytes([0x00, 0xFF, 0xFF, 0x01]) + bytes(raw)
