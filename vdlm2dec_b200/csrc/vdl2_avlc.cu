/* vdl2_avlc.cu -- rows f1/f4 behind the block pipeline: frames -> field records, and the compact hand-over of frames.
 *
 * vdl2_avlc_kernel   one warp per frame.  The 32 lanes copy the frame into shared memory with coalesced 16-byte loads; the
 *                    ACARS CRC (outacars.c:222-230, the one long dependent chain of the walk: up to 2000 table steps) is
 *                    linear, so lanes take 64-byte segments from zero state and a 32-step scan with the "64 zero bytes"
 *                    advance table combines them (the FCS scan of vdl2_link.cu, same polynomial, crc.c); lane 0 then walks
 *                    the dozen header fields (vdl2_avlc.cuh) and stores the 48-byte record.
 * vdl2_frame_rank / _scan / _pack   the frames of one drain leave the device ORDERED (completion order: end of the burst,
 *                    then channel, then position in the block -- the order the reference's blk_thread sees them,
 *                    vdlm2.c:189-206) and PACKED: a 32-byte header per frame + the bytes back to back (16-byte aligned)
 *                    instead of fixed 2048-byte records, with their field records in the same order.  Ranking is a
 *                    counting sort by comparison over compact 16-byte keys staged through shared memory (a few thousand
 *                    frames; reading the keys out of the 2 KB records n times cost 0.88 ms, this costs microseconds),
 *                    offsets a single-CTA scan, packing a warp per frame.
 * Algorithmic traffic: 2048 B in per frame record + 48 B record + header + the frame's own bytes out.
 */
#include <cuda_runtime.h>
#include "vdl2_avlc.cuh"
#include "vdl2_link.h"

#define AVLC_WARPS 4

struct AvlcTab {
	uint16_t adv_lo[256], adv_hi[256];	/* CRC state after 64 zero bytes, by low / high byte of the state */
};
__constant__ AvlcTab c_avlc;

/* CRC of the n bytes at t (shared memory), lane-parallel; every lane returns the result */
__device__ __forceinline__ uint32_t avlc_crc_warp(const uint8_t * t, int n)
{
	const int lane = threadIdx.x & 31;
	uint32_t crc = 0;	/* state at the start of the current 2048-byte round */
	for (int base = 0; base < n; base += 2048) {
		const int s0 = base + 64 * lane, s1 = min(s0 + 64, n);
		uint32_t g = 0;
		for (int i = s0; i < s1; i++)
			g = avlc_crc(g, t[i]);
		/* a short last segment: the bytes that are missing behind it are NOT zero bytes of the message, so the advance below
		   must not be applied for them; segments are processed in order and the scan stops at the end of the data */
		const int nseg = min(32, (n - base + 63) / 64);
		for (int l = 0; l < nseg; l++) {
			const uint32_t gl = __shfl_sync(0xffffffffu, g, l);
			const int seglen = min(64, n - (base + 64 * l));
			if (seglen == 64)
				crc = (uint32_t) (c_avlc.adv_lo[crc & 0xff] ^ c_avlc.adv_hi[crc >> 8]) ^ gl;
			else {	/* the tail: advance the running state through the real bytes one by one (at most 63 steps, once) */
				for (int i = 0; i < seglen; i++)
					crc = avlc_crc(crc, t[base + 64 * l + i]);
			}
		}
	}
	return crc;
}

/* out_index: where the record of frame f goes (NULL: f); nframes_dev: number of frames (NULL: nframes_host) */
__global__ void __launch_bounds__(32 * AVLC_WARPS) vdl2_avlc_kernel(const Vdl2FrameRec * __restrict__ frames, int nframes_host,
								     const unsigned *__restrict__ nframes_dev, const int *__restrict__ out_index,
								     Vdl2AvlcRec * __restrict__ recs)
{
	__shared__ uint4 stage[AVLC_WARPS][sizeof(Vdl2FrameRec) / 16];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nframes = nframes_dev ? (int)min(*nframes_dev, (unsigned)nframes_host) : nframes_host;
	for (int f = blockIdx.x * AVLC_WARPS + warp; f < nframes; f += gridDim.x * AVLC_WARPS) {
		const uint4 *src = reinterpret_cast < const uint4 * >(frames + f);
		const Vdl2FrameRec *fr = reinterpret_cast < const Vdl2FrameRec * >(&stage[warp][0]);
		stage[warp][lane] = src[lane];	/* the header and the first 480 bytes: enough to know the length */
		__syncwarp();
		int l = fr->len;
		l = l < 0 ? 0 : (l > (int)sizeof fr->hdata ? (int)sizeof fr->hdata : l);
		for (int i = lane + 32; i < (32 + l + 15) / 16; i += 32)
			stage[warp][i] = src[i];
		__syncwarp();
		uint32_t crc = 0;
		if (avlc_is_acars(fr->hdata, l))	/* warp uniform */
			crc = avlc_crc_warp(fr->hdata + 13, l - 16 - 1);
		if (lane == 0)
			avlc_fields(fr->hdata, l, crc, recs + (out_index ? out_index[f] : f));
		__syncwarp();
	}
}

/* ---- ordered, packed hand-over of the frames of one drain ---- */
struct Vdl2FrameHdr {		/* identical layout to vdl2_frame_hdr_t (include/vdl2gpu.h) */
	int64_t sync_dump;
	int32_t chn, Fr;
	float ppm;
	int32_t len;
	uint32_t offset;
	int32_t dur;
};
static_assert(sizeof(Vdl2FrameHdr) == 32, "vdl2_frame_hdr_t layout");

/* sort keys out of the 2 KB records into a compact array: the ranking below reads every key n times */
__global__ void vdl2_frame_key_kernel(const Vdl2FrameRec * __restrict__ frames, const unsigned *__restrict__ nframes_dev, unsigned cap,
				      longlong2 * __restrict__ keys)
{
	const int n = (int)min(*nframes_dev, cap);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const Vdl2FrameRec & f = frames[i];
		/* completion order: end of the burst, then channel, then position in the block (length), then arrival */
		keys[i] = make_longlong2(f.sync_dump + f.pad, ((long long)f.chn << 32) | (unsigned)f.len);
	}
}

__global__ void __launch_bounds__(128) vdl2_frame_rank_kernel(const longlong2 * __restrict__ keys, const unsigned *__restrict__ nframes_dev, unsigned cap,
							       int *__restrict__ rank, unsigned *__restrict__ len_sorted)
{
	__shared__ longlong2 tile[128];
	const int n = (int)min(*nframes_dev, cap);
	for (int i0 = blockIdx.x * 128; i0 < n; i0 += gridDim.x * 128) {	/* block-uniform trip count: __syncthreads inside */
		const int i = i0 + (int)threadIdx.x;
		const longlong2 me = i < n ? keys[i] : make_longlong2(0, 0);
		int r = 0;
		for (int base = 0; base < n; base += 128) {
			__syncthreads();
			if (base + (int)threadIdx.x < n)
				tile[threadIdx.x] = keys[base + threadIdx.x];
			__syncthreads();
			const int m = min(128, n - base);
			for (int j = 0; j < m; j++) {
				const longlong2 k = tile[j];
				r += (k.x < me.x) || (k.x == me.x && (k.y < me.y || (k.y == me.y && base + j < i)));
			}
		}
		if (i < n) {
			rank[i] = r;
			len_sorted[r] = (unsigned)((((int)(me.y & 0xffffffffll)) + 15) & ~15);
		}
	}
}

/* exclusive scan of len_sorted (one CTA of 1024 threads); totals[0] = frames, totals[1] = bytes */
__global__ void __launch_bounds__(1024) vdl2_frame_scan_kernel(unsigned *__restrict__ len_sorted, const unsigned *__restrict__ nframes_dev, unsigned cap,
								unsigned *__restrict__ totals)
{
	__shared__ unsigned wsum[32];
	__shared__ unsigned carry_s;
	const int n = (int)min(*nframes_dev, cap);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		carry_s = 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024) {
		const int i = base + threadIdx.x;
		const unsigned v = i < n ? len_sorted[i] : 0u;
		unsigned x = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned t = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o)
				x += t;
		}
		if (lane == 31)
			wsum[warp] = x;
		__syncthreads();
		if (warp == 0) {
			unsigned w = wsum[lane];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
				if (lane >= o)
					w += t;
			}
			wsum[lane] = w;
		}
		__syncthreads();
		const unsigned before = carry_s + (warp ? wsum[warp - 1] : 0u) + x - v;
		if (i < n)
			len_sorted[i] = before;
		__syncthreads();
		if (threadIdx.x == 1023)
			carry_s = before + v;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		totals[0] = (unsigned)n;
		totals[1] = carry_s;
	}
}

__global__ void __launch_bounds__(32 * AVLC_WARPS) vdl2_frame_pack_kernel(const Vdl2FrameRec * __restrict__ frames, const unsigned *__restrict__ nframes_dev,
									   unsigned cap, const int *__restrict__ rank, const unsigned *__restrict__ offs,
									   Vdl2FrameHdr * __restrict__ hdrs, uint8_t * __restrict__ bytes, unsigned bytes_cap)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n = (int)min(*nframes_dev, cap);
	for (int f = blockIdx.x * AVLC_WARPS + warp; f < n; f += gridDim.x * AVLC_WARPS) {
		const Vdl2FrameRec *fr = frames + f;
		const int r = rank[f];
		const unsigned off = offs[r];
		int l = fr->len;
		l = l < 0 ? 0 : (l > (int)sizeof fr->hdata ? (int)sizeof fr->hdata : l);
		if (lane == 0) {
			Vdl2FrameHdr h;
			h.sync_dump = fr->sync_dump;
			h.chn = fr->chn;
			h.Fr = fr->Fr;
			h.ppm = fr->ppm;
			h.len = l;
			h.offset = off;
			h.dur = fr->pad;
			hdrs[r] = h;
		}
		if (off + (unsigned)((l + 15) & ~15) <= bytes_cap) {
			const uint4 *src = reinterpret_cast < const uint4 * >(fr->hdata);	/* hdata sits 32 bytes into a 16-byte aligned record */
			uint4 *dst = reinterpret_cast < uint4 * >(bytes + off);
			for (int i = lane; i < (l + 15) / 16; i += 32)
				dst[i] = src[i];
		}
	}
}

static int upload_tab(void)
{
	static bool done = false;
	if (done)
		return 0;
	static AvlcTab t;
	for (int b = 0; b < 256; b++) {
		uint32_t lo = (uint32_t) b, hi = (uint32_t) b << 8;
		for (int k = 0; k < 64; k++) {
			lo = avlc_crc(lo, 0);
			hi = avlc_crc(hi, 0);
		}
		t.adv_lo[b] = (uint16_t) lo;
		t.adv_hi[b] = (uint16_t) hi;
	}
	const cudaError_t e = cudaMemcpyToSymbol(c_avlc, &t, sizeof t);
	done = (e == cudaSuccess);
	return (int)e;
}

extern "C" int vdl2_avlc_launch(const Vdl2FrameRec * d_frames, int nframes, void *d_recs, void *stream)
{
	if (nframes <= 0)
		return 0;
	if (int e = upload_tab())
		return e;
	const int grid = (nframes + AVLC_WARPS - 1) / AVLC_WARPS;
	vdl2_avlc_kernel <<< grid, 32 * AVLC_WARPS, 0, (cudaStream_t) stream >>> (d_frames, nframes, NULL, NULL, (Vdl2AvlcRec *) d_recs);
	return (int)cudaGetLastError();
}

/* frames (unordered, count on the device) -> rank, offsets, headers + packed bytes, field records in rank order */
extern "C" int vdl2_frames_pack_launch(const Vdl2FrameRec * d_frames, const unsigned *d_nframes, unsigned cap, int *d_rank, unsigned *d_offs,
				       unsigned *d_totals, void *d_hdrs, uint8_t * d_bytes, unsigned bytes_cap, void *d_recs, int expect, void *d_keys,
				       void *stream)
{
	if (int e = upload_tab())
		return e;
	cudaStream_t st = (cudaStream_t) stream;
	const int n = expect > 0 ? expect : 1;
	vdl2_frame_key_kernel <<< (n + 127) / 128, 128, 0, st >>> (d_frames, d_nframes, cap, (longlong2 *) d_keys);
	vdl2_frame_rank_kernel <<< (n + 127) / 128, 128, 0, st >>> ((const longlong2 *)d_keys, d_nframes, cap, d_rank, d_offs);
	vdl2_frame_scan_kernel <<< 1, 1024, 0, st >>> (d_offs, d_nframes, cap, d_totals);
	const int grid = (n + AVLC_WARPS - 1) / AVLC_WARPS;
	vdl2_frame_pack_kernel <<< grid, 32 * AVLC_WARPS, 0, st >>> (d_frames, d_nframes, cap, d_rank, d_offs, (Vdl2FrameHdr *) d_hdrs, d_bytes, bytes_cap);
	if (d_recs)
		vdl2_avlc_kernel <<< grid, 32 * AVLC_WARPS, 0, st >>> (d_frames, (int)cap, d_nframes, d_rank, (Vdl2AvlcRec *) d_recs);
	return (int)cudaGetLastError();
}
