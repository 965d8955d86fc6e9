/*
 * vdl2_link.cu -- the block pipeline behind the demodulator on the GPU.
 *
 * Reference: blk_thread, vdlm2.c:84-161 -- per row rs() (rs.c:81-291, errors-and-erasures RS(255,249)),
 * HDLC bit un-stuffing across the rows (vdlm2.c:116-147), flag framing and check_frame (vdlm2.c:40-61:
 * length >= 13, PPP FCS16 residue 0xf0b8 from crc.c), then out(blk, hdata, l).  The CPU restatement this
 * kernel mirrors step by step is oracle/port/vdl2_link_port.c (pinned against the reference compiled in
 * place); tests/test_gpu_link.py demands bit-identical frames, rs() results and corrected rows.
 *
 * One warp per completed block (<= 8 rows x 255 bytes), everything in shared memory:
 *   RS        syndromes with lanes over the 255 symbols (log/antilog tables in shared memory, XOR
 *             reduction by shuffles); Berlekamp-Massey redundantly on every lane (6 steps, warp
 *             uniform); Chien search with lanes over the 255 positions; Forney with a lane per root.
 *   un-stuff  "a zero after exactly five ones" is a local pattern of the input: every lane takes one
 *             32-bit word (+ 6 bits of the previous one), marks the stuffed zeros with shifts and ANDs,
 *             squeezes them out, and ORs its word into the output at the bit offset given by a warp scan.
 *   framing   the reference's quirks are kept: bytes are OR-ed into hdata[0] until the accumulated value
 *             is a flag; flags right behind it are skipped; the byte index is never reset, so every later
 *             flag closes a candidate frame that starts at hdata[1].
 *   FCS       CRC-16 is linear: lanes take 64-byte segments from zero state, the states at the segment
 *             starts follow from a 32-step scan with a "64 zero bytes" advance table, then every lane
 *             walks its segment once more and tests the residue in front of each flag.
 * This is integer / byte work with table look-ups; its bound is the LSU (shared-memory look-ups), not HBM:
 * algorithmic traffic is 2080 B in + <= 2064 B out per block.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "vdl2_common.h"
#include "vdl2_link.h"

namespace vdl2link {

#define NR 6
#define FCR 120
#define WARPS_PER_CTA 4
#define MAX_GOOD 32

struct Tables {			/* built on the host (vdl2_link_tables), copied to shared memory by every CTA */
	uint8_t gexp[512];
	uint8_t glog[256];
	uint16_t fcs[256];
	uint16_t adv_lo[256], adv_hi[256];	/* FCS state after 64 zero bytes, by low / high byte of the state */
};

__device__ Tables g_tab;

struct WarpMem {
	uint32_t raw[512];	/* the block's data[8][255] as loaded (2040 bytes) */
	uint32_t inw[512];	/* data bytes of the rows back to back (<= 1992), zero padded */
	uint32_t out[512];	/* un-stuffed bit stream */
	uint16_t good[MAX_GOOD];	/* closing-flag positions (relative to the frame start) of frames that passed */
	int root[NR];
};

__device__ __forceinline__ uint8_t gmul(const Tables & T, uint8_t a, uint8_t b)
{
	return (a && b) ? T.gexp[T.glog[a] + T.glog[b]] : (uint8_t) 0;
}

__device__ __forceinline__ uint8_t gpow(const Tables & T, int e)
{				/* alpha^e, e >= 0 */
	return T.gexp[e % 255];
}

__device__ __forceinline__ uint8_t gdiv(const Tables & T, uint8_t a, uint8_t b)
{
	return a ? T.gexp[T.glog[a] + 255 - T.glog[b]] : (uint8_t) 0;
}

/* rs.c:81-291 / port_rs(): one row, all lanes enter; returns the number of corrected symbols or -1 */
__device__ int rs_row(const Tables & T, uint8_t * row, const int *eras, int neras, int *root_sh)
{
	const int lane = threadIdx.x & 31;
	/* syndromes S_i = sum_j row[j] alpha^((FCR+i)(254-j)), lanes over j */
	uint32_t pa = 0, pb = 0;
	for (int j = lane; j < 255; j += 32) {
		const uint8_t d = row[j];
		if (d) {
			const int ld = T.glog[d];
#pragma unroll
			for (int i = 0; i < NR; i++) {
				const uint32_t t = T.gexp[ld + ((FCR + i) * (254 - j)) % 255];
				if (i < 4)
					pa ^= t << (8 * i);
				else
					pb ^= t << (8 * (i - 4));
			}
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		pa ^= __shfl_xor_sync(0xffffffffu, pa, o);
		pb ^= __shfl_xor_sync(0xffffffffu, pb, o);
	}
	if ((pa | pb) == 0)
		return 0;
	uint8_t S[NR], lam[NR + 1], B[NR + 1], Tt[NR + 1], om[NR];
#pragma unroll
	for (int i = 0; i < NR; i++)
		S[i] = (uint8_t) ((i < 4 ? pa >> (8 * i) : pb >> (8 * (i - 4))) & 0xff);
#pragma unroll
	for (int i = 0; i <= NR; i++)
		lam[i] = 0;
	lam[0] = 1;
	for (int e = 0; e < neras; e++) {
		const uint8_t X = gpow(T, 254 - eras[e]);
		for (int j = e + 1; j > 0; j--)
			lam[j] ^= gmul(T, lam[j - 1], X);
	}
#pragma unroll
	for (int i = 0; i <= NR; i++)
		B[i] = lam[i];
	int L = neras;
	for (int r = neras + 1; r <= NR; r++) {
		uint8_t d = 0;
		for (int i = 0; i < r; i++)
			d ^= gmul(T, lam[i], S[r - i - 1]);
		if (d == 0) {
#pragma unroll
			for (int i = NR; i > 0; i--)
				B[i] = B[i - 1];
			B[0] = 0;
			continue;
		}
		Tt[0] = lam[0];
#pragma unroll
		for (int i = 0; i < NR; i++)
			Tt[i + 1] = lam[i + 1] ^ gmul(T, d, B[i]);
		if (2 * L <= r + neras - 1) {
			L = r + neras - L;
#pragma unroll
			for (int i = 0; i <= NR; i++)
				B[i] = gdiv(T, lam[i], d);
		} else {
#pragma unroll
			for (int i = NR; i > 0; i--)
				B[i] = B[i - 1];
			B[0] = 0;
		}
#pragma unroll
		for (int i = 0; i <= NR; i++)
			lam[i] = Tt[i];
	}
	int deg = 0;
#pragma unroll
	for (int i = 0; i <= NR; i++)
		if (lam[i])
			deg = i;
	/* Chien search, lanes over i = 1..255 in ascending order */
	int count = 0;
	for (int base = 1; base <= 255 && count < deg; base += 32) {
		const int i = base + lane;
		uint8_t q = 1;
		if (i <= 255) {
			for (int j = 1; j <= deg; j++)
				if (lam[j])
					q ^= T.gexp[T.glog[lam[j]] + (i * j) % 255];
		}
		unsigned m = __ballot_sync(0xffffffffu, q == 0);
		while (m && count < deg) {
			const int l = __ffs(m) - 1;
			m &= m - 1;
			if (lane == 0)
				root_sh[count] = base + l;
			count++;
		}
	}
	__syncwarp();
	if (count != deg)
		return -1;
	int dom = 0;
#pragma unroll
	for (int i = 0; i < NR; i++) {
		uint8_t t = 0;
		for (int j = (deg < i ? deg : i); j >= 0; j--)
			t ^= gmul(T, S[i - j], lam[j]);
		om[i] = t;
		if (t)
			dom = i;
	}
	/* Forney, a lane per root; the reference walks the roots last to first and gives up at a zero denominator */
	uint8_t num = 0, den = 1;
	int rt = 0;
	if (lane < count) {
		rt = root_sh[lane];
		den = 0;
		for (int i = dom; i >= 0; i--)
			num ^= gmul(T, om[i], gpow(T, i * rt));
		const int top = (deg < NR - 1 ? deg : NR - 1) & ~1;
		for (int i = top; i >= 0; i -= 2)
			den ^= gmul(T, lam[i + 1], gpow(T, i * rt));
	}
	const unsigned zm = __ballot_sync(0xffffffffu, lane < count && den == 0);
	const int jfail = zm ? 31 - __clz(zm) : -1;	/* corrections of roots above it were already applied */
	if (lane < count && lane > jfail && num)
		row[rt - 1] ^= gdiv(T, gmul(T, num, gpow(T, rt * (FCR - 1))), den);
	__syncwarp();
	return zm ? -1 : count;
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
vdl2_link_kernel(const Vdl2BlockRec * __restrict__ blocks, int nblocks, Vdl2FrameRec * frames, unsigned *nframes, unsigned cap,
		 Vdl2BlkStat * stats, uint8_t * rows_after)
{
	__shared__ Tables T;
	__shared__ WarpMem wm[WARPS_PER_CTA];
	{
		const uint32_t *src = reinterpret_cast < const uint32_t * >(&g_tab);
		uint32_t *dst = reinterpret_cast < uint32_t * >(&T);
		for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x)
			dst[i] = src[i];
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int b = blockIdx.x * WARPS_PER_CTA + w;
	if (b >= nblocks)
		return;
	WarpMem & M = wm[w];
	const Vdl2BlockRec *blk = blocks + b;
	const uint32_t *src = reinterpret_cast < const uint32_t * >(blk->data);	/* 8-byte aligned, 510 words */
	for (int i = lane; i < 512; i += 32) {
		M.raw[i] = i < 510 ? src[i] : 0u;
		M.inw[i] = 0u;
		M.out[i] = 0u;
	}
	int nbrow = blk->nbrow, nlbyte = blk->nlbyte;
	nbrow = nbrow < 0 ? 0 : (nbrow > 8 ? 8 : nbrow);
	nlbyte = nlbyte < 0 ? 0 : (nlbyte > 249 ? 249 : nlbyte);
	__syncwarp();
	uint8_t *rawb = reinterpret_cast < uint8_t * >(M.raw);

	/* ---- Reed-Solomon, row by row (vdlm2.c:100-113) ---- */
	int8_t rsres[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	for (int r = 0; r < nbrow; r++) {
		int eras[4] = { 251, 252, 253, 254 }, ne = 0;
		if (r == nbrow - 1) {	/* set_eras, vdlm2.c:63-82 */
			if (nlbyte <= 30)
				ne = 4;
			else if (nlbyte <= 67) {
				ne = 2;
				eras[0] = 253;
				eras[1] = 254;
			}
		}
		rsres[r] = (int8_t) rs_row(T, rawb + 255 * r, eras, ne, M.root);
	}
	if (rows_after)
		for (int i = lane; i < 2040; i += 32)
			rows_after[(size_t) b * 2040 + i] = rawb[i];

	/* ---- the data bytes of the rows back to back ---- */
	const int nbytes = nbrow ? 249 * (nbrow - 1) + nlbyte : 0;
	uint8_t *inb = reinterpret_cast < uint8_t * >(M.inw);
	for (int q = lane; q < nbytes; q += 32) {
		const int r = q / 249, c = q - 249 * r;
		inb[q] = rawb[255 * r + c];
	}
	__syncwarp();

	/* ---- HDLC bit un-stuffing (vdlm2.c:116-131): drop every zero that follows exactly five ones ---- */
	const int nwords = (nbytes + 3) >> 2;
	int obits = 0;		/* kept bits so far (warp uniform) */
	for (int w0 = 0; w0 < nwords; w0 += 32) {
		const int wi = w0 + lane;
		uint32_t cur = 0, prev = 0;
		int vb = 0;
		if (wi < nwords) {
			cur = M.inw[wi];
			prev = wi ? M.inw[wi - 1] : 0u;
			vb = 8 * nbytes - 32 * wi;
			vb = vb > 32 ? 32 : vb;
		}
		const unsigned long long x = ((unsigned long long)cur << 6) | (prev >> 26);
		const unsigned long long c = ~x & (x << 1) & (x << 2) & (x << 3) & (x << 4) & (x << 5) & ~(x << 6);
		const uint32_t valid = vb >= 32 ? 0xffffffffu : ((1u << vb) - 1u);
		unsigned long long rm = (c >> 6) & valid;
		const int nk = vb - __popcll(rm);
		unsigned long long cw = cur & valid;
		while (rm) {	/* squeeze the stuffed zeros out, lowest first (at most 5 per word) */
			const int p = __ffsll((long long)rm) - 1;
			cw = (cw & ((1ull << p) - 1ull)) | ((cw >> (p + 1)) << p);
			rm = (rm >> (p + 1)) << p;
		}
		int off = nk;	/* inclusive scan of the kept-bit counts */
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, off, o);
			if (lane >= o)
				off += t;
		}
		const int total = __shfl_sync(0xffffffffu, off, 31);
		const int pos = obits + off - nk;
		if (nk > 0) {
			const unsigned long long sh = cw << (pos & 31);
			atomicOr(&M.out[pos >> 5], (uint32_t) sh);
			if ((pos & 31) + nk > 32)
				atomicOr(&M.out[(pos >> 5) + 1], (uint32_t) (sh >> 32));
		}
		obits += total;
		__syncwarp();
	}
	const int nob = obits >> 3;	/* complete bytes */
	const uint8_t *ob = reinterpret_cast < const uint8_t * >(M.out);

	/* ---- framing (vdlm2.c:132-147) ---- */
	int n0 = -1;		/* byte at which the OR-accumulated hdata[0] becomes a flag */
	{
		uint32_t carry = 0;
		for (int base = 0; base < nob && n0 < 0; base += 32) {
			const int n = base + lane;
			uint32_t v = carry | (n < nob ? ob[n] : 0u);
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
				if (lane >= o)
					v |= t;
			}
			const unsigned mm = __ballot_sync(0xffffffffu, n < nob && v == 0x7eu);
			if (mm)
				n0 = base + __ffs(mm) - 1;
			carry = __shfl_sync(0xffffffffu, v, 31);
			if (carry & ~0x7eu)
				break;	/* a bit outside the flag pattern: hdata[0] can never become 0x7e */
		}
	}
	int n1 = -1;		/* first byte after n0 that is not a flag: it becomes hdata[1] */
	if (n0 >= 0) {
		for (int base = n0 + 1; base < nob && n1 < 0; base += 32) {
			const int n = base + lane;
			const unsigned mm = __ballot_sync(0xffffffffu, n < nob && ob[n] != 0x7e);
			if (mm)
				n1 = base + __ffs(mm) - 1;
		}
	}
	int kend = n0 < 0 ? 0 : (n1 < 0 ? 1 : nob - n1 + 1), ngood = 0, ntotal = 0;
	if (n1 >= 0) {
		/* ---- FCS16 in front of every later flag (check_frame, vdlm2.c:40-61) ---- */
		const int Mb = nob - n1;	/* bytes from hdata[1] on */
		const int s0 = 64 * lane, s1 = min(s0 + 64, Mb);
		uint32_t g = 0;
		for (int t = s0; t < s1; t++)
			g = (g >> 8) ^ T.fcs[(g ^ ob[n1 + t]) & 0xff];
		uint32_t st = 0xffffu, mine = 0xffffu;
		for (int l = 0; l < 32; l++) {	/* state at the start of every segment */
			if (l == lane)
				mine = st;
			const uint32_t gl = __shfl_sync(0xffffffffu, g, l);
			st = (uint32_t) (T.adv_lo[st & 0xff] ^ T.adv_hi[st >> 8]) ^ gl;
		}
		uint32_t crc = mine;
		unsigned long long hit = 0;	/* bit t: the flag at segment offset t closes a good frame */
		for (int t = s0; t < s1; t++) {
			const uint32_t byte = ob[n1 + t];
			if (byte == 0x7e && t + 2 >= 13 && crc == 0xf0b8u)
				hit |= 1ull << (t - s0);
			crc = (crc >> 8) ^ T.fcs[(crc ^ byte) & 0xff];
		}
		int cnt = __popcll(hit), pre = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, pre, o);
			if (lane >= o)
				pre += t;
		}
		ntotal = __shfl_sync(0xffffffffu, pre, 31);
		int slot = pre - cnt;
		while (hit) {
			const int t = __ffsll((long long)hit) - 1;
			hit &= hit - 1;
			if (slot < MAX_GOOD)
				M.good[slot] = (uint16_t) (s0 + t);
			slot++;
		}
		__syncwarp();
		ngood = ntotal < MAX_GOOD ? ntotal : MAX_GOOD;
		if (lane == 0 && ntotal > MAX_GOOD)
			atomicAdd(nframes + 1, (unsigned)(ntotal - MAX_GOOD));	/* reported by the host as an error: nothing is dropped silently */
	}
	/* ---- hand the frames over: hdata[0] = flag, hdata[1..] = ob[n1..], l = closing flag position + 2 ---- */
	for (int f = 0; f < ngood; f++) {
		const int m = M.good[f], l = m + 2;
		unsigned slot = 0;
		if (lane == 0)
			slot = atomicAdd(nframes, 1u);
		slot = __shfl_sync(0xffffffffu, slot, 0);
		if (slot >= cap)
			continue;	/* counted; the host reports the overflow */
		Vdl2FrameRec *fr = frames + slot;
		if (lane == 0) {
			fr->block = b;
			fr->len = l;
			fr->chn = blk->chn;
			fr->Fr = blk->Fr;
			fr->ppm = blk->ppm;
			fr->pad = (int32_t) (blk->end_dump - blk->sync_dump);	/* burst duration in dumps: completion order for the packed drain */
			fr->sync_dump = blk->sync_dump;
		}
		for (int i = lane; i < l; i += 32)
			fr->hdata[i] = i ? ob[n1 + i - 1] : (uint8_t) 0x7e;
	}
	if (stats && lane == 0) {
		Vdl2BlkStat s;
#pragma unroll
		for (int r = 0; r < 8; r++)
			s.rs[r] = rsres[r];
		s.nbytes = kend;
		s.nframes = ntotal;
		stats[b] = s;
	}
}

}				/* namespace vdl2link */

/* ------------------------------------------------------------------ host shims */
extern "C" int vdl2_link_upload_tables(void)
{
	static vdl2link::Tables t;
	int x = 1;
	for (int i = 0; i < 255; i++) {
		t.gexp[i] = t.gexp[i + 255] = (uint8_t) x;
		t.glog[x] = (uint8_t) i;
		x <<= 1;
		if (x & 0x100)
			x ^= 0x187;	/* x^8 + x^7 + x^2 + x + 1, the field of rs.c */
	}
	t.gexp[510] = t.gexp[0];
	t.gexp[511] = t.gexp[1];
	t.glog[0] = 255;
	for (int b = 0; b < 256; b++) {
		unsigned c = (unsigned)b;
		for (int k = 0; k < 8; k++)
			c = (c & 1u) ? (c >> 1) ^ 0x8408u : c >> 1;
		t.fcs[b] = (uint16_t) c;
	}
	for (int b = 0; b < 256; b++) {	/* the state update is linear over GF(2): advance the two bytes of the state separately */
		unsigned lo = (unsigned)b, hi = (unsigned)b << 8;
		for (int k = 0; k < 64; k++) {
			lo = (lo >> 8) ^ t.fcs[lo & 0xff];
			hi = (hi >> 8) ^ t.fcs[hi & 0xff];
		}
		t.adv_lo[b] = (uint16_t) lo;
		t.adv_hi[b] = (uint16_t) hi;
	}
	return (int)cudaMemcpyToSymbol(vdl2link::g_tab, &t, sizeof t);
}

extern "C" int vdl2_link_launch(const Vdl2BlockRec * d_blocks, int nblocks, Vdl2FrameRec * d_frames, unsigned *d_nframes, unsigned cap,
				Vdl2BlkStat * d_stats, uint8_t * d_rows_after, void *stream)
{
	if (nblocks <= 0)
		return 0;
	const int grid = (nblocks + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	vdl2link::vdl2_link_kernel <<< grid, 32 * WARPS_PER_CTA, 0, (cudaStream_t) stream >>> (d_blocks, nblocks, d_frames, d_nframes, cap, d_stats,
											     d_rows_after);
	return (int)cudaGetLastError();
}
