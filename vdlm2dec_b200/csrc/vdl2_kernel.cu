/*
 * vdl2_kernel.cu -- the fused sm_100a front-end kernel (channeliser + D8PSK demodulator).
 *
 * One persistent single-warp CTA per resident slot; work items are (tile, channel) pairs
 * handed out tile-major by an atomic ticket.  A tile is 32 rows of 1 ms of one channel's
 * input stream (one row per lane), because 1 ms is the joint period of the reference's
 * 21/SDRCLK dump clock and its 25 kHz-periodic NCO table (d8psk.c:348-381, SURVEY.md
 * appendix A.1): every lane then runs the SAME dump schedule with the SAME oscillator
 * value at every step, so the schedule lives in constant memory, the NCO table is a
 * warp-broadcast shared-memory read, and the tap dot products need no cross-lane traffic.
 *
 *   phase 1 (channeliser, rtl.c:285-292 + d8psk.c:366-381): a 3-D TMA tensor map views a
 *     stream as [stream][row][row bytes]; `cp.async.bulk.tensor` boxes of 32 rows x 128 B
 *     land 128B-swizzled in a 3-stage mbarrier ring, which transposes time-major HBM into
 *     lane-major shared memory: lane r reads 16-byte chunk j of its row at (j ^ (r&7))<<4,
 *     conflict free.  Bytes are widened with PRMT magic-number tricks (exact), the complex
 *     MAC runs as packed FFMA2 on sample pairs; dumps go, in time order, to a per-warp scratch that
 *     stays L2 resident (21.6 KB per warp), which keeps shared memory per warp at 13 KB -> 16 warps/SM.
 *     8-bit input at 2 Msps (the RTL path) takes the integer dot-product mixer instead: per-dump TMA
 *     windows, IDP.4A with three weight digits, no conversion (mix_rows_dp4a below).
 *   phase 2 (demodulator): vdl2_demod.cuh, lanes over consecutive steps / symbols.
 *
 * Phase 1 needs no channel state, so a warp mixes tile t of a channel while another warp
 * still demodulates tile t-1; only phase 2 waits on the per-channel progress flag, and while it
 * would wait it runs pass A of the idle search speculatively (stage 0 of the two-stage loop in the
 * kernel, IdlePre in vdl2_demod.cuh).  Consecutive launches may overlap (programmatic dependent
 * launch): progress counts tiles since create, the work counters rotate, scratch slots are per SM.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stddef.h>
#include "vdl2_demod.cuh"
#include "vdl2_kernel.h"

namespace vdl2 {

#define NSTAGE VDL2_NSTAGE
#define STAGE_BYTES 4096

/* ------------------------------------------------------------------ PTX helpers */
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;"::"r" (bar), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r" (bar), "r"(bytes):"memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile ("{\n"
		      ".reg .pred p;\n"
		      "WAIT_%=:\n"
		      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		      "@p bra DONE_%=;\n" "bra WAIT_%=;\n" "DONE_%=:\n" "}"::"r" (bar), "r"(parity):"memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap * map, uint32_t bar, int c0, int c1, int c2,
					    unsigned long long policy)
{
	asm volatile ("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
		      " [%0], [%1, {%3, %4, %5}], [%2], %6;"::"r" (dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
		      :"memory");
}

/* TMA prefetch of a box into L2 (no shared memory, no barrier; SASS UTMAPF): would raise the bytes in flight per warp beyond
   what the stage ring holds.  Tried in round 2 (MM_PREFETCH) on the theory that 16 one-warp CTAs per SM with a few KB in flight
   each cannot cover the DRAM latency -- measured 19-24 % SLOWER at every distance, so it is compiled out. */
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap * map, int c0, int c1, int c2)
{
	asm volatile ("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"::"l" (map), "r"(c0), "r"(c1), "r"(c2):"memory");
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
	unsigned long long ra = *reinterpret_cast < unsigned long long *>(&a);
	unsigned long long rb = *reinterpret_cast < unsigned long long *>(&b);
	unsigned long long rc = *reinterpret_cast < unsigned long long *>(&c);
	unsigned long long rd;
	asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(rd):"l"(ra), "l"(rb), "l"(rc));
	return *reinterpret_cast < float2 * >(&rd);
}

__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
	unsigned long long ra = *reinterpret_cast < unsigned long long *>(&a);
	unsigned long long rb = *reinterpret_cast < unsigned long long *>(&b);
	unsigned long long rd;
	asm("add.rn.f32x2 %0, %1, %2;":"=l"(rd):"l"(ra), "l"(rb));
	return *reinterpret_cast < float2 * >(&rd);
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
	unsigned long long ra = *reinterpret_cast < unsigned long long *>(&a);
	unsigned long long rb = *reinterpret_cast < unsigned long long *>(&b);
	unsigned long long rd;
	asm("mul.rn.f32x2 %0, %1, %2;":"=l"(rd):"l"(ra), "l"(rb));
	return *reinterpret_cast < float2 * >(&rd);
}

#ifndef VDL2_MAX_RESPEC
#define VDL2_MAX_RESPEC 8	/* repeats of the speculative stage per tile (one per burst header decoded while it waits) */
#endif
#ifndef VDL2_CHAIN_POLL_NS
#define VDL2_CHAIN_POLL_NS 100	/* back-off of the wait for the previous tile of the channel */
#endif
#ifndef VDL2_SPEC_AHEAD
#define VDL2_SPEC_AHEAD 12	/* tiles: how far behind the start of a tile the forecast point may lie for the tile to speculate on it */
#endif
/* What a tile starting at dump_base can do ahead of its turn in the channel's chain, from the forecast
   (Vdl2ChanState.fc_dump / fc_clk: the end of the last idle tile, or of the burst whose header was decoded last):
     0..7   the channel is idle there: the tick clock the tile begins with (speculative pass A, see IdlePre);
     8..39  the tile starts inside that burst: 8 + 4 * (first symbol dump mod 8) + tap phase (burst phases ahead, see BurstPre);
     -1     nothing.
   Read without synchronisation (one 16-byte load, written as three words): a torn or stale forecast gives a wrong key, and the
   demodulator verifies whatever was precomputed under it. */
/* the forecast as one aligned 16-byte volatile load: (fc_dump lo, fc_dump hi, fc_clk, fc_df) */
static __device__ __forceinline__ uint4 forecast_load(const Vdl2ChanState * gs)
{
	uint4 v;
	asm volatile ("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];":"=r" (v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w):"l"(&gs->fc_dump));
	return v;
}

static __device__ __forceinline__ int forecast_key(const uint4 fc, long long dump_base, int nd, int &bd0, int &bdlast, float &bdf)
{
	const long long F = (long long)(((unsigned long long)fc.y << 32) | fc.x);
	const int c = (int)fc.z & 7;
	if (F > dump_base) {	/* the last symbol of the burst is at dump F - 1, one symbol every 8 dumps */
		const long long rel = F - 1 - dump_base;
		if (c >= 4 || rel < 24)
			return -1;
		bd0 = (int)(rel & 7);
		bdlast = rel < (long long)nd ? (int)rel : nd - 1;
		bdf = __uint_as_float(fc.w);
		return 8 + 4 * bd0 + c;
	}
	if (dump_base - F > (long long)VDL2_SPEC_AHEAD * VDL2_TILE_DUMPS)
		return -1;	/* too far ahead of the chain to be worth a guess: every burst in between would void it (the tile asks again while it waits) */
	const int s = (c >= 4) ? 0 : 1;	/* dump F + s is the first idle step; then every 2nd dump */
	const long long e = dump_base - F - s;
	return (c & 3) + ((e >= 0 && (e & 1LL) == 0) ? 4 : 0);
}

/* ------------------------------------------------------------------ phase 1: channeliser */
struct MixAcc {			/* partial sums of one dump; .x/.y = even/odd sample of a pair (8-bit formats) */
	float2 A, B, C, G;	/* A: xr*wr  B: xi*wi  C: xr*wi  G: xi*wr   (cf32: A = (I,Q)*re, C = (I,Q)*im) */
};

__device__ __forceinline__ void acc_zero(MixAcc & a)
{
	a.A = a.B = a.C = a.G = make_float2(0.f, 0.f);
}

/* close a dump: D = sum / nf (d8psk.c:377) written to the warp's scratch in time order.
   dc = (1/nf, 1/nf, -cre/nf, -cim/nf): the cu8 path mixes the integers u-127 and removes the
   remaining 0.37f offset of rtl.c:287-289 here, as delta * sum(w) over the dump's oscillator
   values (per channel, per dump, exact to one rounding; zero for the other formats). */
template < int FMT > __device__ __forceinline__ void dump_close(MixAcc & a, float2 * sdrow, const float4 * dcorr, int &k)
{
	const float4 dc = __ldg(dcorr + k);
	float re, im;
	if (FMT == VDL2_FMT_CF32) {
		re = a.A.x - a.C.y;
		im = a.C.x + a.A.y;
	} else if (FMT == VDL2_FMT_F32REAL) {
		re = a.A.x + a.A.y;
		im = a.C.x + a.C.y;
	} else {
		re = (a.A.x + a.A.y) - (a.B.x + a.B.y);
		im = (a.C.x + a.C.y) + (a.G.x + a.G.y);
	}
	__stcg(sdrow + k, ffma2(make_float2(re, im), make_float2(dc.x, dc.y), make_float2(dc.z, dc.w)));
	k++;
	acc_zero(a);
}

/* 4 bytes (I0 Q0 I1 Q1) -> exact floats.  PRMT builds 0x4B0000uu = 2^23 + u and ONE packed add
   removes the bias: cu8 -> u - 127 (the 0.37f remainder is applied per dump, see dump_close),
   cs8 -> s (bytes biased by 128 through the sign-bit flip). */
template < int FMT > __device__ __forceinline__ void cvt_pair8(uint32_t w, float2 & xr, float2 & xi)
{
	if (FMT == VDL2_FMT_CS8)
		w ^= 0x80808080u;
	const uint32_t magic = 0x4B000000u;
	xr.x = __uint_as_float(__byte_perm(w, magic, 0x7440));
	xi.x = __uint_as_float(__byte_perm(w, magic, 0x7441));
	xr.y = __uint_as_float(__byte_perm(w, magic, 0x7442));
	xi.y = __uint_as_float(__byte_perm(w, magic, 0x7443));
	const float bias = (FMT == VDL2_FMT_CS8) ? -8388736.f : -8388735.f;	/* -(2^23 + 128) / -(2^23 + 127) */
	const float2 m = make_float2(bias, bias);
	xr = fadd2(xr, m);
	xi = fadd2(xi, m);
}

/* one sample pair with its oscillator pair W = (re[n], re[n+1], im[n], im[n+1]).
   SPLIT = 0: both samples in the current dump; 1: dump boundary between them; 2: after them */
template < int FMT, int SPLIT > __device__ __forceinline__ void mac_pair(MixAcc & a, float2 xr, float2 xi, float4 W, float2 * sdrow,
									const float4 * dcorr, int &k)
{
	if (SPLIT == 1) {
		a.A.x = fmaf(xr.x, W.x, a.A.x);
		a.B.x = fmaf(xi.x, W.z, a.B.x);
		a.C.x = fmaf(xr.x, W.z, a.C.x);
		a.G.x = fmaf(xi.x, W.x, a.G.x);
		dump_close < FMT > (a, sdrow, dcorr, k);
		a.A.y = xr.y * W.y;
		a.B.y = xi.y * W.w;
		a.C.y = xr.y * W.w;
		a.G.y = xi.y * W.y;
	} else {
		const float2 wr = make_float2(W.x, W.y), wi = make_float2(W.z, W.w);
		a.A = ffma2(xr, wr, a.A);
		a.B = ffma2(xi, wi, a.B);
		a.C = ffma2(xr, wi, a.C);
		a.G = ffma2(xi, wr, a.G);
		if (SPLIT == 2)
			dump_close < FMT > (a, sdrow, dcorr, k);
	}
}

/* a 16-byte chunk of 8-bit IQ = 8 samples = 4 pairs; E = last sample of the current dump
   inside this chunk (0..7), or 8 if the dump continues through the whole chunk */
struct Pairs8 {
	float2 xr[4], xi[4];
	float4 W[4];
};

template < int FMT > __device__ __forceinline__ void load_pairs8(Pairs8 & P, uint4 v, const float4 * w)
{
	const uint32_t d[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
	for (int p = 0; p < 4; p++) {
		cvt_pair8 < FMT > (d[p], P.xr[p], P.xi[p]);
		P.W[p] = w[p];
	}
}

template < int FMT, int E > __device__ __forceinline__ void mac_pairs8(MixAcc & a, const Pairs8 & P, float2 * sdrow, const float4 * dcorr,
									int &k)
{
#pragma unroll
	for (int p = 0; p < 4; p++) {
		if (E == 2 * p)
			mac_pair < FMT, 1 > (a, P.xr[p], P.xi[p], P.W[p], sdrow, dcorr, k);
		else if (E == 2 * p + 1)
			mac_pair < FMT, 2 > (a, P.xr[p], P.xi[p], P.W[p], sdrow, dcorr, k);
		else
			mac_pair < FMT, 0 > (a, P.xr[p], P.xi[p], P.W[p], sdrow, dcorr, k);
	}
}

template < int FMT, int E > __device__ __forceinline__ void chunk8(MixAcc & a, uint4 v, const float4 * w, float2 * sdrow,
								    const float4 * dcorr, int &k)
{
	Pairs8 P;
	load_pairs8 < FMT > (P, v, w);
	mac_pairs8 < FMT, E > (a, P, sdrow, dcorr, k);
}

/* complex float input (the reference's Cbuff, vdlm2.h:89): a chunk is 2 samples; the oscillator
   table is stored duplicated, W = (re, re, im, im) per sample, so that the natural (I, Q) register
   pair feeds FFMA2 directly: A += (I,Q)*(re,re), C += (I,Q)*(im,im); D = (A.x - C.y) + i (C.x + A.y) */
template < int E > __device__ __forceinline__ void chunk_cf32(MixAcc & a, uint4 v, const float4 * w, float2 * sdrow,
							       const float4 * dcorr, int &k)
{
	const float2 x0 = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
	const float2 x1 = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
	const float4 W0 = w[0], W1 = w[1];
	a.A = ffma2(x0, make_float2(W0.x, W0.y), a.A);
	a.C = ffma2(x0, make_float2(W0.z, W0.w), a.C);
	if (E == 0)
		dump_close < VDL2_FMT_CF32 > (a, sdrow, dcorr, k);
	a.A = ffma2(x1, make_float2(W1.x, W1.y), a.A);
	a.C = ffma2(x1, make_float2(W1.z, W1.w), a.C);
	if (E == 1)
		dump_close < VDL2_FMT_CF32 > (a, sdrow, dcorr, k);
}

/* signed 16-bit IQ: a chunk is 4 samples = 2 pairs.  Sign bit flipped, PRMT builds 2^23 + u16,
   one packed add removes 2^23 + 32768: exact. */
__device__ __forceinline__ void cvt_pair16(uint32_t w0, uint32_t w1, float2 & xr, float2 & xi)
{
	w0 ^= 0x80008000u;
	w1 ^= 0x80008000u;
	const uint32_t magic = 0x4B000000u;
	xr.x = __uint_as_float(__byte_perm(w0, magic, 0x7410));
	xi.x = __uint_as_float(__byte_perm(w0, magic, 0x7432));
	xr.y = __uint_as_float(__byte_perm(w1, magic, 0x7410));
	xi.y = __uint_as_float(__byte_perm(w1, magic, 0x7432));
	const float2 m = make_float2(-8421376.f, -8421376.f);	/* -(2^23 + 2^15) */
	xr = fadd2(xr, m);
	xi = fadd2(xi, m);
}

/* real samples (Airspy, air.c:206-208 + d8psk.c:368 with a float Cbuff): D += x * w */
template < int SPLIT > __device__ __forceinline__ void mac_pair_real(MixAcc & a, float2 x, float4 W, float2 * sdrow, const float4 * dcorr,
								      int &k)
{
	if (SPLIT == 1) {
		a.A.x = fmaf(x.x, W.x, a.A.x);
		a.C.x = fmaf(x.x, W.z, a.C.x);
		dump_close < VDL2_FMT_F32REAL > (a, sdrow, dcorr, k);
		a.A.y = x.y * W.y;
		a.C.y = x.y * W.w;
	} else {
		a.A = ffma2(x, make_float2(W.x, W.y), a.A);
		a.C = ffma2(x, make_float2(W.z, W.w), a.C);
		if (SPLIT == 2)
			dump_close < VDL2_FMT_F32REAL > (a, sdrow, dcorr, k);
	}
}

/* a 16-byte chunk of 4 samples (cs16 IQ or float32 real) = 2 pairs; E as in chunk8, 4 = none */
template < int FMT, int E > __device__ __forceinline__ void chunk4(MixAcc & a, uint4 v, const float4 * w, float2 * sdrow,
								    const float4 * dcorr, int &k)
{
	const uint32_t d[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
	for (int p = 0; p < 2; p++) {
		const float4 W = w[p];
		if (FMT == VDL2_FMT_F32REAL) {
			const float2 x = make_float2(__uint_as_float(d[2 * p]), __uint_as_float(d[2 * p + 1]));
			if (E == 2 * p)
				mac_pair_real < 1 > (a, x, W, sdrow, dcorr, k);
			else if (E == 2 * p + 1)
				mac_pair_real < 2 > (a, x, W, sdrow, dcorr, k);
			else
				mac_pair_real < 0 > (a, x, W, sdrow, dcorr, k);
		} else {
			float2 xr, xi;
			cvt_pair16(d[2 * p], d[2 * p + 1], xr, xi);
			if (E == 2 * p)
				mac_pair < FMT, 1 > (a, xr, xi, W, sdrow, dcorr, k);
			else if (E == 2 * p + 1)
				mac_pair < FMT, 2 > (a, xr, xi, W, sdrow, dcorr, k);
			else
				mac_pair < FMT, 0 > (a, xr, xi, W, sdrow, dcorr, k);
		}
	}
}

template < int FMT > struct FmtTraits;
template <> struct FmtTraits <VDL2_FMT_CU8 > { static constexpr int wper_chunk = 4, spc = 8; };
template <> struct FmtTraits <VDL2_FMT_CS8 > { static constexpr int wper_chunk = 4, spc = 8; };
template <> struct FmtTraits <VDL2_FMT_CF32 > { static constexpr int wper_chunk = 2, spc = 2; };
template <> struct FmtTraits <VDL2_FMT_CS16 > { static constexpr int wper_chunk = 2, spc = 4; };
template <> struct FmtTraits <VDL2_FMT_F32REAL > { static constexpr int wper_chunk = 2, spc = 4; };

/* a chunk the current dump runs straight through */
template < int FMT > __device__ __forceinline__ void chunk_plain(MixAcc & a, uint4 v, const float4 * w, float2 * sdrow,
								  const float4 * dcorr, int &k)
{
	if (FMT == VDL2_FMT_CF32)
		chunk_cf32 < 2 > (a, v, w, sdrow, dcorr, k);
	else if (FMT == VDL2_FMT_CS16 || FMT == VDL2_FMT_F32REAL)
		chunk4 < FMT, 4 > (a, v, w, sdrow, dcorr, k);
	else
		chunk8 < FMT, 8 > (a, v, w, sdrow, dcorr, k);
}

/* the chunk in which the current dump ends after sample E */
template < int FMT > __device__ __forceinline__ void chunk_bound(int E, MixAcc & a, uint4 v, const float4 * w, float2 * sdrow,
								  const float4 * dcorr, int &k)
{
	if (FMT == VDL2_FMT_CF32) {
		if (E == 0)
			chunk_cf32 < 0 > (a, v, w, sdrow, dcorr, k);
		else
			chunk_cf32 < 1 > (a, v, w, sdrow, dcorr, k);
	} else if (FMT == VDL2_FMT_CS16 || FMT == VDL2_FMT_F32REAL) {
		switch (E) {
		case 0: chunk4 < FMT, 0 > (a, v, w, sdrow, dcorr, k); break;
		case 1: chunk4 < FMT, 1 > (a, v, w, sdrow, dcorr, k); break;
		case 2: chunk4 < FMT, 2 > (a, v, w, sdrow, dcorr, k); break;
		default: chunk4 < FMT, 3 > (a, v, w, sdrow, dcorr, k); break;
		}
	} else {
		/* conversions and oscillator loads are common to the eight boundary positions */
		Pairs8 P;
		load_pairs8 < FMT > (P, v, w);
		switch (E) {
		case 0: mac_pairs8 < FMT, 0 > (a, P, sdrow, dcorr, k); break;
		case 1: mac_pairs8 < FMT, 1 > (a, P, sdrow, dcorr, k); break;
		case 2: mac_pairs8 < FMT, 2 > (a, P, sdrow, dcorr, k); break;
		case 3: mac_pairs8 < FMT, 3 > (a, P, sdrow, dcorr, k); break;
		case 4: mac_pairs8 < FMT, 4 > (a, P, sdrow, dcorr, k); break;
		case 5: mac_pairs8 < FMT, 5 > (a, P, sdrow, dcorr, k); break;
		case 6: mac_pairs8 < FMT, 6 > (a, P, sdrow, dcorr, k); break;
		default: mac_pairs8 < FMT, 7 > (a, P, sdrow, dcorr, k); break;
		}
	}
}

/* ------------------------------------------------------------------ phase 1, 8-bit formats at 2 Msps: integer mixer
 *
 * Measured on B200 (tools/ubench): byte -> float conversion (2 PRMT + 1 FADD2 per sample) costs more
 * pipe time than the complex MAC itself, and PRMT does not overlap FFMA2.  For cu8/cs8 input the mixer
 * therefore runs on the integer dot-product unit, with no conversion at all:
 *   - the oscillator value w[n] (the reference's float, d8psk.c:353-357) is quantised to 2^-22 and
 *     split into three balanced base-256 digits; one table word per digit holds the four signed bytes
 *     (wr[n], wi[n], wr[n+1], wi[n+1]);
 *   - a data word (I0 Q0 I1 Q1) is made signed with one XOR; a second XOR pattern turns Q into ~Q = -Q-1
 *     so that ONE weight word serves both components:
 *         re += dp4a((I, ~Q), (wr, wi))      im += dp4a((Q, I), (wr, wi))          (IDP.4A, 3 digits each)
 *     the -sum(wi) left by ~Q and the 127.37f / 128 offset of rtl.c:287-289 are linear in w and are
 *     applied per dump from the dcorr table (built in double on the host);
 *   - the six int32 sums are exact; they are combined in fp32 at the dump close (relative error 1e-7,
 *     the weight quantisation contributes < 1e-7 of the dump rms: tests/test_gpu_parity.py).
 * Dump boundaries never fall inside a dot product because the TMA boxes follow the DUMPS: the tensor map
 * has 2-byte elements (one IQ sample); box k = the 32 samples x 32 rows starting at the 8-sample (16-byte,
 * the TMA alignment rule) boundary at or before the first sample of dump k, 64B-swizzled so that LDS.128 is
 * conflict free.  The dump begins 0..7 samples into its window: the word part of that offset selects one of
 * four unrolled bodies (registers cannot be indexed dynamically), the half-word part is a PRMT selector.
 * A 23-sample dump reads one sample too many; its last table entry has zero weights for it.
 * Dumps are transposed through a small shared tile so that the scratch stores are coalesced (one
 * STG.64 per dump used to touch 32 sectors).
 */
#define D8_NST VDL2_D8_NST
#define D8_STAGE 2048		/* 32 rows x 64 bytes */
#define D8_TPITCH 9		/* float2 per row of the transpose tile: 8 dumps + 1 pad (conflict-free STS.64) */

__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c)
{
	return __dp4a((int)a, (int)b, c);
}

/* One dump of one row: d[0..15] = the 64-byte window holding it, the dump starts WO words + (2 bytes if the
   selectors say so) into the window.  WO is a template parameter because registers cannot be indexed
   dynamically; the half-word shift and the (I,Q) -> (Q,I) swap are one PRMT each with run-time selectors. */
template < int FMT, int WO > __device__ __forceinline__ void dump_dp4a(const uint32_t(&d)[16], uint32_t sel_e, uint32_t sel_s,
									 const uint4 * wt, const uint4 * wlast, int (&acc)[6])
{
	const uint32_t KR = (FMT == VDL2_FMT_CU8) ? 0x7F807F80u : 0xFF00FF00u;	/* I -> signed, Q -> ~signed */
	const uint32_t KS = (FMT == VDL2_FMT_CU8) ? 0x80808080u : 0u;
#pragma unroll
	for (int p = 0; p < 12; p++) {
		const uint4 W = (p == 11) ? *wlast : wt[2 * p];
		const uint32_t xr = __byte_perm(d[WO + p], d[WO + p + 1], sel_e) ^ KR;
		const uint32_t xs = __byte_perm(d[WO + p], d[WO + p + 1], sel_s) ^ KS;
		acc[0] = dp4a_ss(xr, W.x, acc[0]);
		acc[3] = dp4a_ss(xs, W.x, acc[3]);
		acc[1] = dp4a_ss(xr, W.y, acc[1]);
		acc[4] = dp4a_ss(xs, W.y, acc[4]);
		acc[2] = dp4a_ss(xr, W.z, acc[2]);
		acc[5] = dp4a_ss(xs, W.z, acc[5]);
	}
}

template < int FMT > __device__ __forceinline__ void mix_rows_dp4a(const CUtensorMap * tmap, const Vdl2KParams & kp, unsigned char *stage0,
								 unsigned long long *bars, float2 * tile, const uint4 * w8, uint32_t & phases,
								 int row0, int stream, const float4 * dcorr, float2 * sd,
								 unsigned long long l2pol)
{
	const int lane = threadIdx.x;
	const unsigned *sched = c_tab.sched_slots[kp.sched_slot];
	/* 64B-swizzled box: 16-byte chunk c of row r sits at r*64 + ((c ^ ((r >> 1) & 3)) << 4) */
	const uint32_t sw = (uint32_t) (lane >> 1) & 3u;
	const unsigned char *rowp = stage0 + lane * 64;
	const uint32_t o0 = (0u ^ sw) << 4, o1 = (1u ^ sw) << 4, o2 = (2u ^ sw) << 4, o3 = (3u ^ sw) << 4;
	int st = 0;
	float4 dcn = __ldg(dcorr);
#pragma unroll 1
	for (int dk = 0; dk < VDL2_DUMPS_PER_ROW; dk++) {
		const float4 dc = dcn;
		const unsigned sk = sched[dk];
		dcn = __ldg(dcorr + (dk + 1 < VDL2_DUMPS_PER_ROW ? dk + 1 : dk));
		mbar_wait(smem_u32(bars + st), (phases >> st) & 1u);
		phases ^= 1u << st;
		const unsigned char *p = rowp + st * D8_STAGE;
		const uint4 v0 = *reinterpret_cast < const uint4 * >(p + o0);
		const uint4 v1 = *reinterpret_cast < const uint4 * >(p + o1);
		const uint4 v2 = *reinterpret_cast < const uint4 * >(p + o2);
		const uint4 v3 = *reinterpret_cast < const uint4 * >(p + o3);
		const uint32_t d[16] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w };
		const unsigned o = (sk >> 16) & 7u;	/* first sample of the dump inside the window */
		const uint32_t sel_e = (o & 1u) ? 0x5432u : 0x3210u, sel_s = (o & 1u) ? 0x4523u : 0x2301u;
		const uint4 *wt = w8 + (sk & 255u);
		const uint4 *wlast = w8 + ((sk >> 8) & 255u);
		int acc[6] = { 0, 0, 0, 0, 0, 0 };
		switch (o >> 1) {
		case 0: dump_dp4a < FMT, 0 > (d, sel_e, sel_s, wt, wlast, acc); break;
		case 1: dump_dp4a < FMT, 1 > (d, sel_e, sel_s, wt, wlast, acc); break;
		case 2: dump_dp4a < FMT, 2 > (d, sel_e, sel_s, wt, wlast, acc); break;
		default: dump_dp4a < FMT, 3 > (d, sel_e, sel_s, wt, wlast, acc); break;
		}
		const float fr = fmaf((float)acc[0], 65536.f, fmaf((float)acc[1], 256.f, (float)acc[2]));
		const float fi = fmaf((float)acc[3], 65536.f, fmaf((float)acc[4], 256.f, (float)acc[5]));
		tile[lane * D8_TPITCH + (dk & 7)] = ffma2(make_float2(fr, fi), make_float2(dc.x, dc.y), make_float2(dc.z, dc.w));
		__syncwarp();	/* every lane has consumed the stage */
		if (lane == 0 && dk + D8_NST < VDL2_DUMPS_PER_ROW) {
			const uint32_t bar = smem_u32(bars + st);
			mbar_expect_tx(bar, D8_STAGE);
			tma_load_3d(smem_u32(stage0 + st * D8_STAGE), tmap, bar, (int)((sched[dk + D8_NST] >> 16) & ~7u), row0, stream, l2pol);
		}
		st = (st + 1 == D8_NST) ? 0 : st + 1;
		if ((dk & 7) == 7 || dk == VDL2_DUMPS_PER_ROW - 1) {
			/* 8 (last group: 4) dumps x 32 rows -> scratch, 64 contiguous bytes per row */
			const int k0 = dk & ~7, ng = dk - k0 + 1;
			const int col = lane & 7, rsub = lane >> 3;
			float2 *dst = sd + VDL2_HIST + k0 + col;
#pragma unroll
			for (int it = 0; it < 8; it++) {
				const int r = it * 4 + rsub;
				if (col < ng)
					__stcg(dst + r * VDL2_DUMPS_PER_ROW, tile[r * D8_TPITCH + col]);
			}
			__syncwarp();
		}
	}
}

/* ------------------------------------------------------------------ phase 1, 8-bit formats at 2 Msps: int8 tensor-core mixer
 *
 * With one row per lane the mixer is a dense integer contraction (vdl2_mma_tables.h has the algebra): per dump
 *     C[32 rows x 8] = A[32 rows x 64 window bytes] * B[64 x 8]      4 x mma.sync.m16n8k32 (u8 or s8 data, s8 weights, int32 sums)
 * instead of 72 IDP.4A + 24 PRMT + 24 LOP3 per row: the raw I,Q bytes ARE the A operand (no conversion, not even a sign flip:
 * unsigned bytes are undone exactly by the accumulator's start value), re and im get their own weight columns (digits 2,1,0 of
 * the oscillator value quantised to 1/8355711), and the accumulator starts at the bits of 1.5 * 2^23 so that it leaves the tensor
 * core as a float.  Same exact int32 sums as the IDP.4A mixer, so parity is untouched.
 *   - HBM -> shared memory: NON-overlapping TMA boxes of 32 rows x 64 bytes (64B swizzle) in a ring; a dump's window is the four
 *     16-byte chunks j0..j0+3 wherever they lie in the ring (at most two boxes), addressed chunk by chunk through ldmatrix.x4
 *     (conflict free under the swizzle).  Every input byte crosses L2 -> SM exactly once (the per-dump windows of the IDP.4A
 *     mixer overlapped: 1.14x DRAM traffic).
 *   - B: per channel one table of 10 window phases (the NCO period is 80 samples = 10 chunks) x 6 columns in fragment order,
 *     one LDS.128 per lane and dump; the samples of the window that belong to the neighbouring dumps are masked with three
 *     shifts (first/last sample of the dump from the schedule word).
 *   - epilogue: lane (g, t) holds columns 2t, 2t+1 of rows g, g+8 (+16 for the second m16 tile).  t even: digits 2 and 1,
 *     t odd: digit 0 (and a duplicate column times zero); one shuffle with lane^1 completes the value, even lanes keep the
 *     first m16 tile and odd lanes the second; t < 2 is the real part, t >= 2 the imaginary part, so a second shuffle with
 *     lane^2 leaves every lane with (re, im) of ONE row.  Four dumps later the lane stores one full 32-byte sector of the
 *     time-ordered scratch straight from registers (no transpose tile: its 2.3 KB pay for a fourth ring stage).
 */
#define MM_NST VDL2_MM_NST
#define MM_STAGE 2048		/* 32 rows x 64 bytes */
#ifndef MM_PREFETCH
#define MM_PREFETCH 0		/* boxes prefetched into L2 ahead of the stage ring: A/B on B200 (profiles/r2_ab_prefetch.txt): 0 -> 2.69 ms, 8 -> 3.20, 16 -> 3.26, 24 -> 3.34 ms per step: off */
#endif
#ifndef MM_QUAD_STORE
#define MM_QUAD_STORE 1		/* dumps leave the registers a quad of dumps at a time, one dump per lane of a quad (see the exchange in mix_rows_mma): a
				   store instruction touches 8 rows with a full 32-byte sector each instead of 32 rows with half a sector each.  0: A/B */
#endif
#ifndef MM_UNROLL
#define MM_UNROLL 2		/* dumps per store group: 2 = half a 32-byte sector per lane and store (A/B on B200: 1 % faster than 4, smaller loop) */
#endif

template < int FMT > __device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t(&a)[4], uint32_t b0, uint32_t b1)
{
	if (FMT == VDL2_FMT_CU8)
		asm volatile ("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};":"+r" (c[0]),
			      "+r"(c[1]), "+r"(c[2]), "+r"(c[3]):"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
	else
		asm volatile ("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};":"+r" (c[0]),
			      "+r"(c[1]), "+r"(c[2]), "+r"(c[3]):"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

/* first k step of a dump: D = A B + (cx, cy, cx, cy), the accumulator start values straight from the table registers (no copies) */
template < int FMT > __device__ __forceinline__ void imma_16832_init(int (&d)[4], const uint32_t(&a)[4], uint32_t b0, uint32_t b1, int cx, int cy)
{
	if (FMT == VDL2_FMT_CU8)
		asm volatile ("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};":"=r" (d[0]),
			      "=r"(d[1]), "=r"(d[2]), "=r"(d[3]):"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(cx), "r"(cy));
	else
		asm volatile ("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};":"=r" (d[0]),
			      "=r"(d[1]), "=r"(d[2]), "=r"(d[3]):"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(cx), "r"(cy));
}

__device__ __forceinline__ void ldsm_x4(uint32_t(&a)[4], uint32_t addr)
{
	asm volatile ("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];":"=r" (a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]):"r"(addr));
}

/* shifts with PTX semantics: amounts >= 32 give 0 (the C operators are undefined there) */
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n)
{
	uint32_t r;
	asm("shl.b32 %0, %1, %2;":"=r"(r):"r"(v), "r"(n));
	return r;
}

__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t n)
{
	uint32_t r;
	asm("shr.u32 %0, %1, %2;":"=r"(r):"r"(v), "r"(n));
	return r;
}

/* 16-byte read-only load with an L2 eviction hint: the per-channel tables (9 KB per channel, re-read for every tile of the
   channel) are kept with evict_last -- without it the streaming input pushes them out and they come back from DRAM
   (0.63 GB of 9.3 GB read per step in the round-2 v15 profile) */
__device__ __forceinline__ uint4 ldg_keep(const void *p, unsigned long long policy)
{
	uint4 v;
	asm volatile ("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;":"=r" (v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w):"l"(p), "l"(policy));
	return v;
}

/* one elected lane (the warp is converged): lets ptxas issue the TMA from uniform registers without a broadcast loop */
__device__ __forceinline__ bool elect_one()
{
	uint32_t p;
	asm volatile ("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}":"=r" (p));
	return p != 0;
}

/* loop invariants the compiler would otherwise recompute inside the dump loop (it rematerialises cheap per-lane values at
   128 registers): an empty asm makes the value opaque, so it stays in its register */
#define VDL2_PIN(v) asm volatile ("" : "+r"(v))
#define VDL2_PINF(v) asm volatile ("" : "+f"(v))

template < int FMT > __device__ __forceinline__ void mix_rows_mma(const CUtensorMap * tmap, const Vdl2KParams & kp, unsigned char *stage0,
								unsigned long long *bars, const uint4 * bt, uint32_t & phases, int row0, int stream,
								const int4 * dtab, float2 * sd, unsigned long long l2pol, unsigned long long l2keep)
{
	const int lane = threadIdx.x;
	const unsigned *sched = c_tab.sched_slots[kp.sched_slot];
	const int g = lane >> 2, t = lane & 3;
	/* ldmatrix.x4: lane l supplies row l & 7 of matrix l >> 3; matrix i = rows 8 (i & 1) .. +7 of an m16 tile, window chunk
	   2 s + (i >> 1).  64B-swizzled box: chunk c of row r sits at r * 64 + ((c ^ ((r >> 1) & 3)) << 4).  Chunk numbers are
	   kept multiplied by 16 (byte offsets). */
	uint32_t sw16 = ((uint32_t) (lane >> 1) & 3u) << 4, cwl16 = ((uint32_t) lane >> 4) << 4;
	uint32_t rowoff = smem_u32(stage0) + (((uint32_t) lane >> 3) & 1u) * 512u + ((uint32_t) lane & 7u) * 64u;
	uint32_t btl = smem_u32(bt + ((g >> 2) * 3 + min(g & 3, 2)) * 4 + t);	/* columns 3 and 7 re-read 2 and 5: finite, multiplied by 0 */
	const int4 *dp = dtab + t;
	float scx = (t & 1) ? 1.f : 65536.f, scy = (t & 1) ? 0.f : 256.f;
	float nscx = -12582912.f * scx, nscy = -12582912.f * scy;	/* exact: powers of two */
	int t32 = 32 * t;
	uint32_t odd = (uint32_t) t & 1u, hi = (uint32_t) t >> 1;
	/* after the two exchanges below this lane owns (re, im) of ONE row: g + 16 (t & 1) + 8 (t >> 1); four consecutive dumps of it
	   are one 32-byte sector of the scratch */
#if MM_QUAD_STORE
	/* ... or, with the exchange over a quad of dumps, dump 4 q + t of the FOUR rows g, g + 8, g + 16, g + 24 */
	float2 *dstq = sd + VDL2_HIST + g * VDL2_DUMPS_PER_ROW + t;
	float keep[4] = { 0.f, 0.f, 0.f, 0.f };	/* first pair of the quad: this lane's component of the dump it keeps, rows g + 8 j */
	static_assert(MM_UNROLL == 2 && VDL2_DUMPS_PER_ROW % 4 == 0, "the quad exchange runs over two pairs of dumps");
#else
	float4 *dst = reinterpret_cast < float4 * >(sd + VDL2_HIST + (g + 16 * (t & 1) + 8 * (t >> 1)) * VDL2_DUMPS_PER_ROW);
#endif
	VDL2_PIN(sw16);
	VDL2_PIN(cwl16);
	VDL2_PIN(rowoff);
	VDL2_PIN(btl);
	VDL2_PINF(scx);
	VDL2_PINF(scy);
	VDL2_PINF(nscx);
	VDL2_PINF(nscy);
	VDL2_PIN(t32);
	VDL2_PIN(odd);
	VDL2_PIN(hi);
	const float2 sc = make_float2(scx, scy), nsc = make_float2(nscx, nscy);
	const int nbox = kp.nbox;
	int st = 0, box = 0;
	uint4 dn = ldg_keep(dp, l2keep);
	mbar_wait(smem_u32(bars), phases & 1u);
	phases ^= 1u;
	static_assert(VDL2_DUMPS_PER_ROW % MM_UNROLL == 0 && MM_NST >= 4, "whole store groups per row; the phase-2 scratch needs four stages of room");
#ifndef MM_PAIR_UNROLL
#define MM_PAIR_UNROLL 1	/* A/B: 2 = both pairs of a quad in one loop body (no copies of the kept values, static quad parity; twice the code) */
#endif
	constexpr int pair_unroll = MM_PAIR_UNROLL;
#pragma unroll pair_unroll
	for (int dk0 = 0; dk0 < VDL2_DUMPS_PER_ROW; dk0 += MM_UNROLL) {
		/* software pipeline over the MM_UNROLL dumps of a store group: first every dump's window and weights go to registers
		   (waits, ldmatrix, B + masks), then the tensor-core products and epilogues run back to back while the loads of the
		   later dumps are still in flight */
		uint32_t a00[MM_UNROLL][4], a01[MM_UNROLL][4], a10[MM_UNROLL][4], a11[MM_UNROLL][4];	/* [dump][m16 tile, k step] */
		uint4 B[MM_UNROLL];
		int4 dc[MM_UNROLL];
		int rst[MM_UNROLL], rbox[MM_UNROLL];	/* stage / box to refill after the dump (-1: none) */
#pragma unroll
		for (int u = 0; u < MM_UNROLL; u++) {
			dc[u] = make_int4((int)dn.x, (int)dn.y, (int)dn.z, (int)dn.w);
			const unsigned sk = sched[dk0 + u];
			dp += 4;
			dn = ldg_keep(dp, l2keep);	/* the table carries one entry more than there are dumps */
			const int st1 = (st + 1 == MM_NST) ? 0 : st + 1;
			if (sk & VDL2_MM_W) {
				mbar_wait(smem_u32(bars + st1), (phases >> st1) & 1u);
				phases ^= 1u << st1;
			}
			const uint32_t base0 = rowoff + (uint32_t) st * MM_STAGE, base1 = rowoff + (uint32_t) st1 * MM_STAGE;
			const uint32_t q0 = (sk & 0x30u) + cwl16, q1 = q0 + 32u;	/* 16 * (chunk of this lane's matrix counted from the box start) */
			const uint32_t ad0 = (q0 >= 64u ? base1 : base0) + ((q0 ^ sw16) & 0x30u);
			const uint32_t ad1 = (q1 >= 64u ? base1 : base0) + ((q1 ^ sw16) & 0x30u);
			ldsm_x4(a00[u], ad0);
			ldsm_x4(a10[u], ad0 + 1024u);
			ldsm_x4(a01[u], ad1);
			ldsm_x4(a11[u], ad1 + 1024u);
			asm volatile ("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];":"=r" (B[u].x), "=r"(B[u].y), "=r"(B[u].z), "=r"(B[u].w):"r"(btl + ((sk & 0x3f00u) >> 2)));
			const int o16 = (int)((sk >> 16) & 127u), te = t32 - (int)(sk >> 23);
			B[u].x &= shl_clamp(0xffffffffu, (uint32_t) max(o16 - t32, 0));	/* samples 2t, 2t+1: keep those >= o */
			B[u].z &= shr_clamp(0xffffffffu, (uint32_t) max(te + 288, 0));	/* samples 16+2t, 17+2t: keep those < e */
			B[u].w &= shr_clamp(0xffffffffu, (uint32_t) max(te + 416, 0));	/* samples 24+2t, 25+2t */
			rst[u] = -1;
			rbox[u] = 0;
			if (sk & VDL2_MM_R) {	/* the next dump starts in the next box: this one can be refilled once its window is in registers */
				if (box + MM_NST < nbox) {
					rst[u] = st;
					rbox[u] = box + MM_NST;
				}
				box++;
				st = st1;
			}
		}
#if MM_QUAD_STORE
		float pp[MM_UNROLL][4];	/* [dump of the pair][row g + 8 j]: this lane's digits of its component, weighed and summed */
#else
		float2 out[MM_UNROLL];
#endif
#pragma unroll
		for (int u = 0; u < MM_UNROLL; u++) {
			int c0[4], c1[4];
			imma_16832_init < FMT > (c0, a00[u], B[u].x, B[u].y, dc[u].x, dc[u].y);
			imma_16832_init < FMT > (c1, a10[u], B[u].x, B[u].y, dc[u].x, dc[u].y);
			imma_16832 < FMT > (c0, a01[u], B[u].z, B[u].w);
			imma_16832 < FMT > (c1, a11[u], B[u].z, B[u].w);
			if (rst[u] >= 0) {	/* the products above consumed the window: ldmatrix has completed for the whole warp */
				__syncwarp();
				if (elect_one()) {
					const uint32_t bar = smem_u32(bars + rst[u]);
					mbar_expect_tx(bar, MM_STAGE);
					tma_load_3d(smem_u32(stage0 + rst[u] * MM_STAGE), tmap, bar, rbox[u] * 32, row0, stream, l2pol);
					if (MM_PREFETCH && rbox[u] + MM_PREFETCH < nbox)
						tma_prefetch_3d(tmap, (rbox[u] + MM_PREFETCH) * 32, row0, stream);
				}
			}
			/* the accumulators are floats 12582912 + sum: remove the bias and weigh the digits in one exact FFMA2 each */
			const float2 y0 = ffma2(make_float2(__int_as_float(c0[0]), __int_as_float(c0[1])), sc, nsc);
			const float2 y1 = ffma2(make_float2(__int_as_float(c0[2]), __int_as_float(c0[3])), sc, nsc);
			const float2 y2 = ffma2(make_float2(__int_as_float(c1[0]), __int_as_float(c1[1])), sc, nsc);
			const float2 y3 = ffma2(make_float2(__int_as_float(c1[2]), __int_as_float(c1[3])), sc, nsc);
			const float p0 = y0.x + y0.y, p1 = y1.x + y1.y, p2 = y2.x + y2.y, p3 = y3.x + y3.y;	/* rows g, g + 8, g + 16, g + 24 */
#if MM_QUAD_STORE
			pp[u][0] = p0;
			pp[u][1] = p1;
			pp[u][2] = p2;
			pp[u][3] = p3;
#else
			/* lane ^ 1 holds the other digits: even lanes complete rows g, g + 8, odd lanes rows g + 16, g + 24 */
			const float r0 = __shfl_xor_sync(0xffffffffu, odd ? p0 : p2, 1);
			const float r1 = __shfl_xor_sync(0xffffffffu, odd ? p1 : p3, 1);
			const float sf = __int_as_float(dc[u].z), corr = __int_as_float(dc[u].w);
			const float v0 = fmaf((odd ? p2 : p0) + r0, sf, corr), v1 = fmaf((odd ? p3 : p1) + r1, sf, corr);
			/* lane ^ 2 holds the other component of the same two rows: t < 2 keeps the first row, t >= 2 the second */
			const float rx = __shfl_xor_sync(0xffffffffu, hi ? v0 : v1, 2);
			out[u] = make_float2(hi ? rx : v0, hi ? v1 : rx);
#endif
		}
#if MM_QUAD_STORE
		/* Exchange over a quad of dumps 4 q .. 4 q + 3 (two trips through this loop): the four lanes of a quad (same g) hold digits
		   (lane ^ 1) and components (lane ^ 2) of the same four rows, and end up with ONE dump each -- lane t with dump 4 q + t -- of
		   all four.  Same number of shuffles as an exchange per dump (3 per dump), same sums in the same order (own + received), but
		   a store instruction then touches 8 rows with one full sector each instead of 32 rows with half a sector each: the scatter
		   of the row-per-lane layout was 15 % of all LSU wavefronts of the kernel (ncu, round 2 v19).
		   round A, per pair: a lane keeps the dump of the pair with its own parity and completes its component of it */
		{
			const float sfk = __int_as_float(odd ? dc[1].z : dc[0].z), corrk = __int_as_float(odd ? dc[1].w : dc[0].w);
			float va[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const float x = __shfl_xor_sync(0xffffffffu, odd ? pp[0][j] : pp[1][j], 1);
				va[j] = fmaf((odd ? pp[1][j] : pp[0][j]) + x, sfk, corrk);
			}
			if (!(dk0 & 2)) {
#pragma unroll
				for (int j = 0; j < 4; j++)
					keep[j] = va[j];
			} else {
				/* round B, per quad: t < 2 (real parts) keeps the first pair's dump, t >= 2 (imaginary parts) the second pair's */
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const float x = __shfl_xor_sync(0xffffffffu, hi ? keep[j] : va[j], 2);
					const float mine = hi ? va[j] : keep[j];
					/* evict-last: the scratch is read back from L2 in phase 2 */
					asm volatile ("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;"::"l" (dstq + j * 8 * VDL2_DUMPS_PER_ROW), "f"(hi ? x : mine),
						      "f"(hi ? mine : x), "l"(l2keep):"memory");
				}
				dstq += 4;
			}
		}
#else
		/* MM_UNROLL dumps of this lane's row, contiguous in the time-ordered scratch; evict-last: it is read back from L2 in phase 2 */
		asm volatile ("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"::"l" (dst), "f"(out[0].x), "f"(out[0].y), "f"(out[1].x), "f"(out[1].y),
			      "l"(l2keep):"memory");
#if MM_UNROLL == 4
		asm volatile ("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"::"l" (dst + 1), "f"(out[2].x), "f"(out[2].y), "f"(out[3].x), "f"(out[3].y),
			      "l"(l2keep):"memory");
#endif
		dst += MM_UNROLL / 2;
#endif
	}
}

/* ------------------------------------------------------------------ the kernel */
template < int FMT, int DP, bool TAPS > __global__ void __launch_bounds__(32, VDL2_MIN_CTAS)
#ifdef VDL2_KP_BYVALUE
vdl2_frontend_kernel(const __grid_constant__ CUtensorMap tmap, const Vdl2KParams kp)
#else
vdl2_frontend_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Vdl2KParams kp)
#endif
{
	/* TAPS = false: production launches (no observation taps, no debug flags) run an instantiation in which all
	   of that code is compiled out -- the kernel has to stay inside the instruction cache */
	extern __shared__ __align__(1024) unsigned char smem[];
	const int lane = threadIdx.x;
	unsigned char *stage0 = smem;
	/* stages (+ the transpose tile of the integer mixer) | mbarriers | header soft bits | oscillator table */
	/* DP: 0 generic fp32 mixer, 1 IDP.4A mixer, 2 int8 tensor-core mixer */
	constexpr int STAGES_BYTES = DP == 2 ? MM_NST * MM_STAGE : (DP ? D8_NST * D8_STAGE + 32 * D8_TPITCH * 8 : NSTAGE * STAGE_BYTES);
	constexpr int NBAR = DP == 2 ? MM_NST : (DP ? D8_NST : NSTAGE);
	unsigned long long *bars = reinterpret_cast < unsigned long long *>(smem + STAGES_BYTES);
	float *hv = reinterpret_cast < float *>(smem + STAGES_BYTES + 64);
	float4 *wsm = reinterpret_cast < float4 * >(hv + 32);
	/* the decimated stream of the tile lives in global memory (L2): sd[0..15] history, sd[16 + i] dump i */
	/* Scratch slot: one per CTA that can be resident on this SM.  Consecutive launches may overlap (programmatic
	   dependent launch: the next launch's CTAs move in while this one's tail drains), so the slot is taken from a
	   per-SM bit mask instead of blockIdx. */
	unsigned smid;
	asm volatile ("mov.u32 %0, %%smid;":"=r" (smid));
	unsigned slotbit = 0;
	if (lane == 0) {
		unsigned *m = kp.slotmask + smid;
		const unsigned all = kp.slots_per_sm >= 32 ? 0xffffffffu : ((1u << kp.slots_per_sm) - 1u);
		for (;;) {
			const unsigned freeb = ~(*(volatile unsigned *)m) & all;
			if (!freeb) {
				__nanosleep(100);
				continue;
			}
			const unsigned bit = freeb & (0u - freeb);
			if (!(atomicOr(m, bit) & bit)) {
				slotbit = bit;
				break;
			}
		}
	}
	slotbit = __shfl_sync(0xffffffffu, slotbit, 0);
	float2 *sd = kp.scratch + ((size_t) smid * kp.slots_per_sm + (__ffs(slotbit) - 1)) * (VDL2_HIST + VDL2_TILE_DUMPS);
	/* At most two launches are active at a time: this one waits until the launch before the previous one has
	   completed (ticket[3] counts completed launches; they complete in order because every launch continues every
	   channel's chain).  Work counters rotate over three slots: this launch uses ticket[ticket_sel] and clears the one
	   the NEXT launch will use (last used two launches ago).  Only then may the next launch start. */
	if (lane == 0) {
		/* counters count since create and wrap after 2^32: compare differences, never values */
		while ((int)((unsigned)kp.launch_seq - 1u - *(volatile unsigned *)(kp.ticket + 3)) > 0)
			__nanosleep(500);
		if (blockIdx.x == 0) {
			kp.ticket[(kp.ticket_sel + 1) % 3] = 0u;
			__threadfence();
		}
	}
	__syncwarp();
	asm volatile ("griddepcontrol.launch_dependents;":::"memory");
	/* phase 2 scratch lives in the TMA stages, which are idle while a tile is demodulated */
	IdleScratch scr;
	scr.pht = reinterpret_cast < float *>(stage0);
	scr.vw = reinterpret_cast < float2 * >(stage0 + VDL2_PHT_LEN * 4);
	scr.win = reinterpret_cast < float2 * >(stage0 + VDL2_PHT_LEN * 4 + 96 * 8);
	scr.cand = reinterpret_cast < unsigned short *>(stage0 + VDL2_PHT_LEN * 4 + 96 * 8 + VDL2_WIN_LEN * 8);
	scr.cand0 = scr.cand + VDL2_CAND_CAP;
	scr.hb = reinterpret_cast < unsigned char *>(scr.cand0 + VDL2_CAND0_CAP);
	/* burst_prephase stages its windows in vw .. cand0: they must be one contiguous run of VDL2_BPRE_BUF float2 (they are, by the lines above) */
	static_assert(VDL2_PHT_LEN * 4 + 96 * 8 + VDL2_WIN_LEN * 8 + (VDL2_CAND_CAP + VDL2_CAND0_CAP) * 2 + VDL2_TILE_DUMPS / 8 <= STAGES_BYTES,
		      "phase 2 scratch must fit the stages");
	static_assert(NBAR <= 8, "mbarriers live in 64 bytes");

	if ((smem_u32(smem) & 1023u) != 0)
		__trap();	/* the 128B swizzle pattern assumes 1 KiB aligned stages */
	if (lane == 0) {
		for (int s = 0; s < NBAR; s++)
			mbar_init(smem_u32(bars + s), 1);
		asm volatile ("fence.mbarrier_init.release.cluster;":::"memory");
	}
	__syncwarp();

	uint32_t phases = 0;	/* parity bit per stage */
	const int nitems = kp.ntiles * kp.nch;
	const int wpc = FmtTraits < FMT >::wper_chunk;
	const int nbox = kp.nbox;
	const int nco = kp.nco_pairs;
	unsigned long long l2pol;	/* the input is read exactly once: evict-first keeps the per-warp scratch L2 resident */
#ifdef VDL2_IN_EVICT_NORMAL	/* A/B */
	asm volatile ("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;":"=l" (l2pol));
#else
	asm volatile ("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;":"=l" (l2pol));
#endif
	unsigned long long l2keep;	/* the per-warp scratch: written in phase 1, read back in phase 2, should never reach DRAM */
#ifdef VDL2_MM_NOKEEP	/* A/B */
	asm volatile ("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;":"=l" (l2keep));
#else
	asm volatile ("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;":"=l" (l2keep));
#endif
	const uint32_t l7 = (uint32_t) (lane & 7);
	float2 *sdrow = sd + VDL2_HIST + VDL2_DUMPS_PER_ROW * lane;

	for (;;) {
		int item = 0;
		if (lane == 0)
			item = (int)atomicAdd(kp.ticket + kp.ticket_sel, 1u);
		item = __shfl_sync(0xffffffffu, item, 0);
		if (item >= nitems)
			break;
		const int tile = item / kp.nch;
		const int ch = item - tile * kp.nch;
		const int stream = ch / kp.ch_per_stream;
		const int row0 = tile * VDL2_ROWS_PER_TILE;
		const int nrows = min(VDL2_ROWS_PER_TILE, kp.nrows - row0);
		const int nd = nrows * VDL2_DUMPS_PER_ROW;

		const float4 *dcorr = kp.dcorr + (size_t) ch * VDL2_DUMPS_PER_ROW;
		/* phase 2 of the previous item used the stage memory through the generic proxy (its scratch aliases the stages):
		   order those accesses before the asynchronous-proxy writes of the TMA loads issued below */
		asm volatile ("fence.proxy.async.shared::cta;":::"memory");
		if (DP == 2) {
			/* ---- phase 1, int8 tensor-core mixer: non-overlapping 64-byte boxes, see mix_rows_mma ---- */
			if (lane == 0) {
				for (int b = 0; b < MM_NST; b++) {
					const uint32_t bar = smem_u32(bars + b);
					mbar_expect_tx(bar, MM_STAGE);
					tma_load_3d(smem_u32(stage0 + b * MM_STAGE), &tmap, bar, b * 32, row0, stream, l2pol);
				}
				for (int b = MM_NST; b < MM_NST + MM_PREFETCH; b++)
					tma_prefetch_3d(&tmap, b * 32, row0, stream);
			}
			uint4 *btsm = reinterpret_cast < uint4 * >(wsm);
			for (int i = lane; i < VDL2_MM_BT_ENTRIES; i += 32)
				btsm[i] = ldg_keep(kp.w8 + (size_t) ch * VDL2_MM_BT_ENTRIES + i, l2keep);
			__syncwarp();
			mix_rows_mma < FMT > (&tmap, kp, stage0, bars, btsm, phases, row0, stream,
					      reinterpret_cast < const int4 * >(kp.dcorr) + (size_t) ch * VDL2_MM_DT_ENTRIES, sd, l2pol, l2keep);
		} else if (DP) {
			/* ---- phase 1, integer mixer: dump-aligned boxes, see mix_rows_dp4a ---- */
			if (lane == 0) {
				for (int b = 0; b < D8_NST; b++) {
					const uint32_t bar = smem_u32(bars + b);
					mbar_expect_tx(bar, D8_STAGE);
					tma_load_3d(smem_u32(stage0 + b * D8_STAGE), &tmap, bar, (int)((c_tab.sched_slots[kp.sched_slot][b] >> 16) & ~7u), row0,
						    stream, l2pol);
				}
			}
			uint4 *w8 = reinterpret_cast < uint4 * >(wsm);
			for (int i = lane; i < VDL2_W8_ENTRIES; i += 32)
				w8[i] = __ldg(kp.w8 + (size_t) ch * VDL2_W8_ENTRIES + i);
			__syncwarp();
			mix_rows_dp4a < FMT > (&tmap, kp, stage0, bars, reinterpret_cast < float2 * >(smem + D8_NST * D8_STAGE), w8, phases, row0,
					       stream, dcorr, sd, l2pol);
		} else {
			/* prologue: first boxes in flight, oscillator table to shared memory */
			if (lane == 0) {
				for (int b = 0; b < NSTAGE && b < nbox; b++) {
					const uint32_t bar = smem_u32(bars + b);
					mbar_expect_tx(bar, STAGE_BYTES);
					tma_load_3d(smem_u32(stage0 + b * STAGE_BYTES), &tmap, bar, b * 32, row0, stream, l2pol);
				}
			}
			for (int i = lane; i < nco + kp.wext; i += 32)
				wsm[i] = kp.wtab[(size_t) ch * nco + (i < nco ? i : i - nco)];
			__syncwarp();

			/* ---- phase 1: dump-centric walk over the row.  Dump k = np plain chunks + the chunk it ends in
			   (after sample E); chunks come from the TMA ring, 8 per 128-byte box; the oscillator table is
			   addressed from a per-dump base (no wrap inside a dump: the table is extended). ---- */
			MixAcc acc;
			acc_zero(acc);
			int k = 0, bx = 0, j = 0;
			mbar_wait(smem_u32(bars), phases & 1u);
			phases ^= 1u;
			const unsigned char *rowp = stage0 + lane * 128;
	#define VDL2_LOAD_CHUNK() (*reinterpret_cast < const uint4 * >(rowp + ((j ^ l7) << 4)))
	#define VDL2_NEXT_CHUNK()                                                                                  \
			do {                                                                                        \
				if (++j == 8) {                                                                     \
					j = 0;                                                                      \
					__syncwarp();                                                               \
					const int slot_ = bx % NSTAGE;                                              \
					if (lane == 0 && bx + NSTAGE < nbox) {                                      \
						const uint32_t bar_ = smem_u32(bars + slot_);                       \
						mbar_expect_tx(bar_, STAGE_BYTES);                                  \
						tma_load_3d(smem_u32(stage0 + slot_ * STAGE_BYTES), &tmap, bar_, (bx + NSTAGE) * 32, row0, stream, l2pol); \
					}                                                                           \
					bx++;                                                                       \
					if (bx < nbox) {                                                            \
						const int ns_ = bx % NSTAGE;                                        \
						mbar_wait(smem_u32(bars + ns_), (phases >> ns_) & 1u);              \
						phases ^= 1u << ns_;                                                \
						rowp = stage0 + ns_ * STAGE_BYTES + lane * 128;                     \
					}                                                                           \
				}                                                                                   \
			} while (0)
	#pragma unroll 1
			for (int dk = 0; dk < VDL2_DUMPS_PER_ROW; dk++) {
				const unsigned sk = c_tab.sched_slots[kp.sched_slot][dk];
				const int E = (int)((sk >> 8) & 255u);
				const float4 *w = wsm + (sk >> 16);
				int np = (int)(sk & 255u);
	#ifdef VDL2_UNROLL_NP2
				if (np == 2) {	/* the common shape at 2 Msps: 24 (23) samples = 2 whole chunks + the boundary chunk */
					uint4 v = VDL2_LOAD_CHUNK();
					chunk_plain < FMT > (acc, v, w, sdrow, dcorr, k);
					VDL2_NEXT_CHUNK();
					v = VDL2_LOAD_CHUNK();
					chunk_plain < FMT > (acc, v, w + wpc, sdrow, dcorr, k);
					VDL2_NEXT_CHUNK();
					w += 2 * wpc;
				} else
	#endif
				{
	#pragma unroll 1
					for (; np > 0; np--) {
						const uint4 v = VDL2_LOAD_CHUNK();
						chunk_plain < FMT > (acc, v, w, sdrow, dcorr, k);
						VDL2_NEXT_CHUNK();
						w += wpc;
					}
				}
				const uint4 v = VDL2_LOAD_CHUNK();
				chunk_bound < FMT > (E, acc, v, w, sdrow, dcorr, k);
				VDL2_NEXT_CHUNK();
			}
	#undef VDL2_NEXT_CHUNK
	#undef VDL2_LOAD_CHUNK
		}
		__threadfence_block();
		__syncwarp();

		/* ---- two stages through ONE instance of the demodulator code:
		   stage 0 (speculative, see IdlePre): while the previous tile of this channel is still being demodulated
		     elsewhere, pass A of the idle search from the channel's published forecast.  Without it the per-channel
		     chain of phase 2 executions is the critical path of the launch;
		   stage 1: wait for the previous tile, load its state, demodulate. ---- */
		Vdl2ChanState *gs = kp.state + ch;
		/* last 16 dumps of the tile, the next tile's filter history: fetched now, not on the chain (an L2 round trip) */
		const float2 htail = __ldcg(sd + nd + (lane & (VDL2_HIST - 1)));
		IdlePre pre;
		pre.valid = 0;
		pre.used = 0;
		ChanRegs R;
		unsigned n_dumps = 0;
		int chn = 0, Fr = 0, nph = 0, had_pre = 0, idle_at_start = 0;
#ifdef VDL2_CHAIN_STATS
		unsigned long long cs_t0 = 0, cs_t1 = 0, cs_ta = 0, cs_tb = 0;
		long long cs_sync0 = 0;
#endif
		int guess = -1, nrespec = 0;	/* what the speculative stage assumed (forecast_key; -1: nothing), times it was repeated */
		BurstPre bp;
		bp.valid = 0;
		const long long dump_base = kp.dump_base + (long long)row0 * VDL2_DUMPS_PER_ROW;
		const bool can_spec = !(TAPS && ((kp.taps & VDL2_TAP_STEPS_BIT) || (kp.flags & (VDL2_FLAG_NO_SCREEN | VDL2_FLAG_NO_PREPASS))));
#pragma unroll 1
		for (int stage = 0; stage < 2; stage++) {
			const bool spec = (stage == 0);
			if (spec) {
				if (!can_spec)
					continue;
				/* the channel's forecast (Vdl2ChanState.fc_*), read without synchronisation: idle_run verifies the guess */
				int bd0 = 0, bdlast = 0;
				float bdf = 0.f;
				/* whatever an earlier pass of this stage left is void: both kinds of results share S.pht */
				bp.valid = 0;
				pre.valid = 0;
				pre.pos0 = 0;
				guess = forecast_key(forecast_load(gs), dump_base, nd, bd0, bdlast, bdf);
				if (guess < 0)
					continue;
				if (guess >= 8) {	/* as far as is known the tile starts inside a burst */
					vdl2::burst_prephase(kp, sd, scr, bp, bd0, bdlast, guess & 3, bdf);
					if (bdlast >= nd - 1)
						continue;
					/* ... that ends inside it, at dump bdlast: the idle search of the rest of the tile, ahead of the chain as well */
					pre.pos0 = bdlast + 1;
				}
				memset(&R, 0, sizeof R);
				R.clk = guess >= 8 ? (guess & 3) : guess;
				R.state = VDL2_ST_WSYNC;
				R.perr = 100.f;
			} else {
				/* wait for the previous tile of this channel, load its state.  While waiting, watch the forecast: a burst
				   header decoded further up the chain changes the tick clock every later tile will start with, and a
				   tile that speculated on the old one (with few channels EVERY later tile of the launch has: they all
				   start at once) would otherwise repeat pass A inside the chain, one tile after the other */
				int redo = 0;
				if (lane == 0) {
					const volatile int *pr = kp.progress + ch;
					(void)pr;
					for (;;) {
						/* progress and forecast in flight together: one L2 round trip per poll, not two */
#ifdef VDL2_CHAIN_ACQ_FENCE	/* A/B: volatile load here, fence.acq_rel.gpu after the loop (it also waits for this lane's own scratch stores) */
						const unsigned done = (unsigned)*pr;
#else
						unsigned done;	/* acquire side of the hand-over: the shuffle below extends it to the other lanes */
						asm volatile ("ld.acquire.gpu.global.u32 %0, [%1];":"=r" (done):"l"(kp.progress + ch):"memory");
#endif
						const uint4 fc = forecast_load(gs);
						if ((int)((unsigned)(kp.tile_base + tile) - done) <= 0)	/* tiles completed since create (wrap-safe): launches may overlap */
							break;
						if (can_spec && nrespec < VDL2_MAX_RESPEC) {
							int t0, t1;
							float t2;
							const int g2 = forecast_key(fc, dump_base, nd, t0, t1, t2);
							if (g2 >= 0 && g2 != guess) {
								redo = 1;
								break;
							}
						}
						__nanosleep(VDL2_CHAIN_POLL_NS);
					}
#if !defined(VDL2_CHAIN_FENCE_SC) && defined(VDL2_CHAIN_ACQ_FENCE)
					/* acquire side of the hand-over: one fence in the polling lane, the warp shuffle below extends it to the others
					   (a sequentially consistent fence in every lane, __threadfence(), costs a microsecond per tile of the chain) */
					asm volatile ("fence.acq_rel.gpu;":::"memory");
#endif
				}
#ifdef VDL2_CHAIN_STATS
				asm volatile ("mov.u64 %0, %%globaltimer;":"=l" (cs_t0));
#endif
				redo = __shfl_sync(0xffffffffu, redo, 0);
				if (redo) {
					nrespec++;
					stage = -1;
					continue;
				}
#ifdef VDL2_CHAIN_FENCE_SC	/* A/B */
				__threadfence();
#endif
				/* every load of the state first, the store of the history last: the compiler cannot move a load across a global store
				   that might alias it, and a store in the middle made this two dependent L2 round trips (1.2 us per tile of the chain) */
				float2 hist01 = make_float2(0.f, 0.f);
				if (lane < VDL2_HIST)
					hist01 = make_float2(__ldcg(gs->hist_re + lane), __ldcg(gs->hist_im + lane));
				scr.pht[lane] = __ldcg(gs->ph + lane);
				scr.pht[lane + 32] = __ldcg(gs->ph + lane + 32);
				if (lane < 28)
					hv[lane] = __ldcg(gs->hv + lane);
				R.perr = __ldcg(&gs->perr);
				R.p2err = __ldcg(&gs->p2err);
				R.pfr = __ldcg(&gs->pfr);
				R.df = __ldcg(&gs->df);
				R.P1 = __ldcg(&gs->P1);
				R.ppm = __ldcg(&gs->ppm);
				R.clk = __ldcg(&gs->clk);
				R.state = __ldcg(&gs->state);
				R.symidx = __ldcg(&gs->symidx);
				R.nbrow = __ldcg(&gs->nbrow);
				R.nlbyte = __ldcg(&gs->nlbyte);
				R.bytes_done = __ldcg(&gs->bytes_done);
				R.bitacc = __ldcg(&gs->bitacc);
				R.nbitacc = __ldcg(&gs->nbitacc);
				R.sync_dump = __ldcg(&gs->sync_dump);
				R.n_steps = __ldcg(&gs->n_steps);
				R.n_syncs = __ldcg(&gs->n_syncs);
				R.n_syms = __ldcg(&gs->n_syms);
#ifdef VDL2_STATE_LOAD_FENCE	/* A/B: round 1 split the state load in two dependent L2 round trips here; __syncwarp below orders the history stores */
				__threadfence_block();
#endif
				n_dumps = __ldcg(&gs->n_dumps);
				chn = __ldcg(&gs->chn);
				Fr = __ldcg(&gs->Fr);
				if (lane < VDL2_HIST)
					__stcg(sd + lane, hist01);
				__syncwarp();
				if (TAPS && (kp.taps & VDL2_TAP_DUMPS_BIT)) {
					float2 *dst = kp.tap_dumps + (size_t) ch * kp.cap_dumps;
					for (int i = lane; i < nd; i += 32)
						if (n_dumps + i < kp.cap_dumps)
							dst[n_dumps + i] = __ldcg(sd + VDL2_HIST + i);
					n_dumps += nd;
				}
				had_pre = pre.valid;
				idle_at_start = (R.state == VDL2_ST_WSYNC);
#ifdef VDL2_CHAIN_STATS
				cs_sync0 = R.sync_dump;
				asm volatile ("mov.u64 %0, %%globaltimer;":"=l" (cs_ta));
#endif
			}
			nph = 0;
			demod_tile < TAPS > (kp, ch, chn, Fr, R, sd, scr, hv, nd, dump_base, nph, pre, spec, bp);
			__syncwarp();
#ifdef VDL2_CHAIN_STATS
			asm volatile ("mov.u64 %0, %%globaltimer;":"=l" (cs_tb));
#endif
		}
#ifndef VDL2_NO_STATS
		if (lane == 0)	/* statistics: 0 speculation used, 1 wasted, 2 idle tile without one, 3 tile that starts inside a burst */
			atomicAdd(kp.ticket + 12 + (pre.used ? 0 : (had_pre ? 1 : (idle_at_start ? 2 : 3))), 1u);
#endif

		/* ---- store state, release the channel ---- */
		if (lane < VDL2_HIST) {
			gs->hist_re[lane] = htail.x;
			gs->hist_im[lane] = htail.y;
		}
		gs->ph[lane] = scr.pht[nph + lane];
		gs->ph[lane + 32] = scr.pht[nph + lane + 32];
		if (lane < 28)
			gs->hv[lane] = hv[lane];
		if (lane == 0) {
			gs->perr = R.perr;
			gs->p2err = R.p2err;
			gs->pfr = R.pfr;
			gs->df = R.df;
			gs->P1 = R.P1;
			gs->ppm = R.ppm;
			gs->clk = R.clk;
			gs->state = R.state;
			gs->symidx = R.symidx;
			gs->nbrow = R.nbrow;
			gs->nlbyte = R.nlbyte;
			gs->bytes_done = R.bytes_done;
			gs->bitacc = R.bitacc;
			gs->nbitacc = R.nbitacc;
			gs->sync_dump = R.sync_dump;
			gs->n_steps = R.n_steps;
			gs->n_syncs = R.n_syncs;
			gs->n_syms = R.n_syms;
			gs->n_dumps = n_dumps;
			if (R.state == VDL2_ST_WSYNC) {
				gs->fc_dump = dump_base + nd;
				gs->fc_clk = R.clk;
			}
		}
#ifdef VDL2_CHAIN_FENCE_SC	/* A/B */
		__threadfence();
		__syncwarp();
#else
		/* release side: every lane's stores are ordered before the barrier, the publishing lane's fence makes them visible
		   before the progress counter moves (the pattern of a grid-wide barrier) */
		__syncwarp();
		if (lane == 0)
			asm volatile ("fence.acq_rel.gpu;":::"memory");
#endif
#ifdef VDL2_CHAIN_STATS	/* debug build: time on the chain (previous tile ready -> this tile published) by kind of tile, ticket[16 + 4 * kind] */
		if (lane == 0) {
			asm volatile ("mov.u64 %0, %%globaltimer;":"=l" (cs_t1));
			const bool idle_end = R.state == VDL2_ST_WSYNC, trig = R.sync_dump != cs_sync0;
			const int kind = idle_at_start ? (trig ? (idle_end ? 3 : 2) : (pre.used ? 0 : 1)) : (idle_end ? (trig ? 6 : 5) : 4);
			atomicAdd(kp.ticket + 16 + 4 * kind, 1u);
			atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 16 + 4 * kind + 2), cs_t1 - cs_t0);
			if (kind == 0) {	/* the plain idle step in three parts: state load, demodulator, state store + fence */
				atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 48), cs_ta - cs_t0);
				atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 50), cs_tb - cs_ta);
				atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 52), cs_t1 - cs_tb);
			}
		}
#endif
		if (lane == 0)
			atomicExch(kp.progress + ch, kp.tile_base + tile + 1);
		__syncwarp();
	}
	if (lane == 0) {
		atomicAnd(kp.slotmask + smid, ~slotbit);
		/* the last CTA out marks the launch as completed */
		if (atomicAdd(kp.ticket + 5 + kp.ticket_sel, 1u) == gridDim.x - 1) {
			kp.ticket[5 + kp.ticket_sel] = 0u;
			__threadfence();
			atomicAdd(kp.ticket + 3, 1u);
		}
	}
}

/* ------------------------------------------------------------------ row f3: wideband shared-stream channeliser
 *
 * One pass over a stream produces the decimated 84 ksps streams of ALL its channels (generalises d8psk.c:353-381, where every
 * channel thread re-reads the whole Cbuff): the raw bytes of a dump window go through ldmatrix ONCE and are multiplied by the
 * weight fragments of every channel in turn.  In tensor-core terms the channel axis is simply more columns of B --
 * C[32 rows x 8 nch] = A[32 x 64] * B[64 x 8 nch] -- i.e. a pruned DFT whose "bins" are the reference's own oscillator
 * tables, so every output is the SAME exact integer sum as in the fused kernel (bit-identical to its T1 tap).
 * Items are (tile, stream); weights and per-dump constants come straight from L2 (they are shared by every warp working on
 * the stream, 9 KB per channel), outputs go to HBM in time order: out[channel][row * 84 + dump].
 */
struct Vdl2ChanlParams {
	int nstreams, cps, nrows, ntiles, nbox, sched_slot;
	const uint4 *bt;	/* [nch][VDL2_MM_BT_ENTRIES] */
	const int4 *dt;		/* [nch][VDL2_MM_DT_ENTRIES] */
	float2 *out;		/* [nch][out_pitch] */
	size_t out_pitch;	/* float2 per channel */
	unsigned *ticket;
};

template < int FMT > __global__ void __launch_bounds__(32, 16) vdl2_channelise_kernel(const __grid_constant__ CUtensorMap tmap, const Vdl2ChanlParams cp)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	const int lane = threadIdx.x;
	unsigned char *stage0 = smem;
	unsigned long long *bars = reinterpret_cast < unsigned long long *>(smem + MM_NST * MM_STAGE);
	if (lane == 0) {
		for (int s = 0; s < MM_NST; s++)
			mbar_init(smem_u32(bars + s), 1);
		asm volatile ("fence.mbarrier_init.release.cluster;":::"memory");
	}
	__syncwarp();
	unsigned long long l2pol;
	asm volatile ("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;":"=l" (l2pol));
	const unsigned *sched = c_tab.sched_slots[cp.sched_slot];
	const int g = lane >> 2, t = lane & 3;
	const uint32_t sw16 = ((uint32_t) (lane >> 1) & 3u) << 4, cwl16 = ((uint32_t) lane >> 4) << 4;
	const uint32_t rowoff = smem_u32(stage0) + (((uint32_t) lane >> 3) & 1u) * 512u + ((uint32_t) lane & 7u) * 64u;
	const int btl = ((g >> 2) * 3 + min(g & 3, 2)) * 4 + t;
	const float2 sc = (t & 1) ? make_float2(1.f, 0.f) : make_float2(65536.f, 256.f);
	const float2 nsc = make_float2(-12582912.f * sc.x, -12582912.f * sc.y);
	const int t32 = 32 * t;
	const bool odd = (t & 1) != 0, hi = (t >> 1) != 0;
	const int rowl = g + 16 * (t & 1) + 8 * (t >> 1);	/* the row this lane owns after the two exchanges */
#ifdef VDL2_CHANL_NOQUAD	/* A/B */
	const bool quad = false;
#else
	const bool quad = (cp.cps == 1);
#endif
	uint32_t phases = 0;
	const int nitems = cp.ntiles * cp.nstreams;
	for (;;) {
		int item = 0;
		if (lane == 0)
			item = (int)atomicAdd(cp.ticket, 1u);
		item = __shfl_sync(0xffffffffu, item, 0);
		if (item >= nitems)
			break;
		const int tile = item / cp.nstreams, stream = item - tile * cp.nstreams;
		const int row0 = tile * VDL2_ROWS_PER_TILE;
		const bool rowok = row0 + rowl < cp.nrows;
		asm volatile ("fence.proxy.async.shared::cta;":::"memory");
		if (lane == 0)
			for (int b = 0; b < MM_NST; b++) {
				const uint32_t bar = smem_u32(bars + b);
				mbar_expect_tx(bar, MM_STAGE);
				tma_load_3d(smem_u32(stage0 + b * MM_STAGE), &tmap, bar, b * 32, row0, stream, l2pol);
			}
		int st = 0, box = 0;
		float keep[4] = { 0.f, 0.f, 0.f, 0.f };	/* quad exchange (one channel per stream): first pair's kept values */
		uint4 Bn[2];
		int4 dcn[2];
#pragma unroll
		for (int u = 0; u < 2; u++) {	/* weights / constants of (pair 0, channel 0) */
			Bn[u] = __ldg(cp.bt + (size_t) (stream * cp.cps) * VDL2_MM_BT_ENTRIES + btl + ((sched[u] & 0x3f00u) >> 6));
			dcn[u] = __ldg(cp.dt + (size_t) (stream * cp.cps) * VDL2_MM_DT_ENTRIES + 4 * u + t);
		}
		mbar_wait(smem_u32(bars), phases & 1u);
		phases ^= 1u;
#pragma unroll 1
		for (int dk0 = 0; dk0 < VDL2_DUMPS_PER_ROW; dk0 += 2) {
			uint32_t a00[2][4], a01[2][4], a10[2][4], a11[2][4];
			unsigned skv[2];
			int rst[2], rbox[2];
#pragma unroll
			for (int u = 0; u < 2; u++) {
				const unsigned sk = sched[dk0 + u];
				skv[u] = sk;
				const int st1 = (st + 1 == MM_NST) ? 0 : st + 1;
				if (sk & VDL2_MM_W) {
					mbar_wait(smem_u32(bars + st1), (phases >> st1) & 1u);
					phases ^= 1u << st1;
				}
				const uint32_t base0 = rowoff + (uint32_t) st * MM_STAGE, base1 = rowoff + (uint32_t) st1 * MM_STAGE;
				const uint32_t q0 = (sk & 0x30u) + cwl16, q1 = q0 + 32u;
				const uint32_t ad0 = (q0 >= 64u ? base1 : base0) + ((q0 ^ sw16) & 0x30u);
				const uint32_t ad1 = (q1 >= 64u ? base1 : base0) + ((q1 ^ sw16) & 0x30u);
				ldsm_x4(a00[u], ad0);
				ldsm_x4(a10[u], ad0 + 1024u);
				ldsm_x4(a01[u], ad1);
				ldsm_x4(a11[u], ad1 + 1024u);
				rst[u] = -1;
				rbox[u] = 0;
				if (sk & VDL2_MM_R) {
					if (box + MM_NST < cp.nbox) {
						rst[u] = st;
						rbox[u] = box + MM_NST;
					}
					box++;
					st = st1;
				}
			}
			/* the windows are in registers: every channel of the stream takes its turn with its own weights.  The weight
			   fragments and per-dump constants come from L2: those of the NEXT (channel, dump pair) are requested before
			   the current ones are used (the first profile of this kernel spent 10 cycles per issued instruction on them) */
#pragma unroll 1
			for (int c = 0; c < cp.cps; c++) {
				const int ch = stream * cp.cps + c;
				uint4 Bc[2];
				int4 dcc[2];
#pragma unroll
				for (int u = 0; u < 2; u++) {
					Bc[u] = Bn[u];
					dcc[u] = dcn[u];
				}
				{	/* next iteration: channel c + 1 of this pair, or channel 0 of the next pair (clamped at the end of the row) */
					const bool wrap = (c + 1 == cp.cps);
					const int cn = wrap ? 0 : c + 1, dkn = wrap ? (dk0 + 2 < VDL2_DUMPS_PER_ROW ? dk0 + 2 : dk0) : dk0;
					const uint4 *btn = cp.bt + (size_t) (stream * cp.cps + cn) * VDL2_MM_BT_ENTRIES + btl;
					const int4 *dtn = cp.dt + (size_t) (stream * cp.cps + cn) * VDL2_MM_DT_ENTRIES + dkn * 4 + t;
#pragma unroll
					for (int u = 0; u < 2; u++) {
						Bn[u] = __ldg(btn + ((sched[dkn + u] & 0x3f00u) >> 6));
						dcn[u] = __ldg(dtn + 4 * u);
					}
				}
				float2 o[2];
				float pp[2][4];
#pragma unroll
				for (int u = 0; u < 2; u++) {
					const unsigned sk = skv[u];
					uint4 B = Bc[u];	/* fragment of window phase 6 p * 4 entries of 16 bytes */
					const int4 dc = dcc[u];
					const int o16 = (int)((sk >> 16) & 127u), te = t32 - (int)(sk >> 23);
					B.x &= shl_clamp(0xffffffffu, (uint32_t) max(o16 - t32, 0));
					B.z &= shr_clamp(0xffffffffu, (uint32_t) max(te + 288, 0));
					B.w &= shr_clamp(0xffffffffu, (uint32_t) max(te + 416, 0));
					int c0[4], c1[4];
					imma_16832_init < FMT > (c0, a00[u], B.x, B.y, dc.x, dc.y);
					imma_16832_init < FMT > (c1, a10[u], B.x, B.y, dc.x, dc.y);
					imma_16832 < FMT > (c0, a01[u], B.z, B.w);
					imma_16832 < FMT > (c1, a11[u], B.z, B.w);
					const float2 y0 = ffma2(make_float2(__int_as_float(c0[0]), __int_as_float(c0[1])), sc, nsc);
					const float2 y1 = ffma2(make_float2(__int_as_float(c0[2]), __int_as_float(c0[3])), sc, nsc);
					const float2 y2 = ffma2(make_float2(__int_as_float(c1[0]), __int_as_float(c1[1])), sc, nsc);
					const float2 y3 = ffma2(make_float2(__int_as_float(c1[2]), __int_as_float(c1[3])), sc, nsc);
					const float p0 = y0.x + y0.y, p1 = y1.x + y1.y, p2 = y2.x + y2.y, p3 = y3.x + y3.y;
					pp[u][0] = p0;
					pp[u][1] = p1;
					pp[u][2] = p2;
					pp[u][3] = p3;
					if (!quad) {
						const float r0 = __shfl_xor_sync(0xffffffffu, odd ? p0 : p2, 1);
						const float r1 = __shfl_xor_sync(0xffffffffu, odd ? p1 : p3, 1);
						const float sf = __int_as_float(dc.z), corr = __int_as_float(dc.w);
						const float v0 = fmaf((odd ? p2 : p0) + r0, sf, corr), v1 = fmaf((odd ? p3 : p1) + r1, sf, corr);
						const float rx = __shfl_xor_sync(0xffffffffu, hi ? v0 : v1, 2);
						o[u] = make_float2(hi ? rx : v0, hi ? v1 : rx);
					}
				}
				if (!quad) {
					if (rowok)
						__stcs(reinterpret_cast < float4 * >(cp.out + (size_t) ch * cp.out_pitch + (size_t) (row0 + rowl) * VDL2_DUMPS_PER_ROW + dk0),
						       make_float4(o[0].x, o[0].y, o[1].x, o[1].y));
				} else {
					/* one channel per stream: the exchange over a quad of dumps of the fused mixer (mix_rows_mma) -- lane t ends up with
					   dump 4 q + t of the rows g, g + 8, g + 16, g + 24, so a store touches 8 rows with a full sector each.  Same sums in
					   the same order: the values stay bit identical to the fused kernel's */
					const float sfk = __int_as_float(odd ? dcc[1].z : dcc[0].z), corrk = __int_as_float(odd ? dcc[1].w : dcc[0].w);
					float va[4];
#pragma unroll
					for (int j = 0; j < 4; j++) {
						const float x = __shfl_xor_sync(0xffffffffu, odd ? pp[0][j] : pp[1][j], 1);
						va[j] = fmaf((odd ? pp[1][j] : pp[0][j]) + x, sfk, corrk);
					}
					if (!(dk0 & 2)) {
#pragma unroll
						for (int j = 0; j < 4; j++)
							keep[j] = va[j];
					} else {
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const float x = __shfl_xor_sync(0xffffffffu, hi ? keep[j] : va[j], 2);
							const float mine = hi ? va[j] : keep[j];
							if (row0 + g + 8 * j < cp.nrows)
								__stcs(cp.out + (size_t) ch * cp.out_pitch + (size_t) (row0 + g + 8 * j) * VDL2_DUMPS_PER_ROW + (dk0 - 2) + t,
								       make_float2(hi ? x : mine, hi ? mine : x));
						}
					}
				}
			}
			__syncwarp();
#pragma unroll
			for (int u = 0; u < 2; u++)
				if (rst[u] >= 0 && elect_one()) {
					const uint32_t bar = smem_u32(bars + rst[u]);
					mbar_expect_tx(bar, MM_STAGE);
					tma_load_3d(smem_u32(stage0 + rst[u] * MM_STAGE), &tmap, bar, rbox[u] * 32, row0, stream, l2pol);
				}
		}
	}
}

/* highest SM id + 1 (ids are not contiguous on parts with disabled SMs): sizes the scratch and the slot masks */
__global__ void vdl2_nsmid_kernel(unsigned *out)
{
	unsigned n;
	asm volatile ("mov.u32 %0, %%nsmid;":"=r" (n));
	*out = n;
}

}				/* namespace vdl2 */

/* ------------------------------------------------------------------ launch shims used by vdl2_host.cu */
extern "C" int vdl2_kernel_smem_bytes(int nco_entries, int dp4a)
{
	if (dp4a == 2)
		return MM_NST * MM_STAGE + 64 + 32 * 4 + VDL2_MM_BT_ENTRIES * 16;
	if (dp4a)
		return D8_NST * D8_STAGE + 32 * D8_TPITCH * 8 + 64 + 32 * 4 + VDL2_W8_ENTRIES * 16;
	return VDL2_NSTAGE * STAGE_BYTES + 64 + 32 * 4 + nco_entries * 16;
}

template < int FMT, int DP, bool TAPS > static cudaError_t launch_fmt2(const CUtensorMap & tmap, const Vdl2KParams & kp, int grid, int smem,
									 cudaStream_t st, int pdl)
{
	cudaError_t e = cudaFuncSetAttribute(vdl2::vdl2_frontend_kernel < FMT, DP, TAPS >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if (e != cudaSuccess)
		return e;
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3((unsigned)grid);
	cfg.blockDim = dim3(32);
	cfg.dynamicSmemBytes = (size_t) smem;
	cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;	/* the next launch may start while this one's tail drains */
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	/* only when the caller asked for overlapping launches AND the operation in front of this one on the stream is a front-end
	   launch: the kernel never executes griddepcontrol.wait, so anything else in front (the rtl.c expansion kernel, a copy into
	   the staging buffer) must have completed in the ordinary way before the first TMA load */
	cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, vdl2::vdl2_frontend_kernel < FMT, DP, TAPS >, tmap, kp);
}

template < int FMT, int DP > static cudaError_t launch_fmt(const CUtensorMap & tmap, const Vdl2KParams & kp, int grid, int smem, cudaStream_t st, int pdl)
{
#ifdef VDL2_ALWAYS_TAPS
	if (true)
#else
	if (kp.taps || kp.flags)
#endif
		return launch_fmt2 < FMT, DP, true > (tmap, kp, grid, smem, st, pdl);
	return launch_fmt2 < FMT, DP, false > (tmap, kp, grid, smem, st, pdl);
}

extern "C" int vdl2_kernel_launch(int fmt, int dp4a, const void *tmap, const Vdl2KParams * kp, int grid, int smem, void *stream, int pdl)
{
	const CUtensorMap & m = *reinterpret_cast < const CUtensorMap * >(tmap);
	cudaStream_t st = (cudaStream_t) stream;
	if (dp4a == 2) {
		switch (fmt) {
		case VDL2_FMT_CU8: return (int)launch_fmt < VDL2_FMT_CU8, 2 > (m, *kp, grid, smem, st, pdl);
		case VDL2_FMT_CS8: return (int)launch_fmt < VDL2_FMT_CS8, 2 > (m, *kp, grid, smem, st, pdl);
		}
		return (int)cudaErrorInvalidValue;
	}
	if (dp4a) {
		switch (fmt) {
		case VDL2_FMT_CU8: return (int)launch_fmt < VDL2_FMT_CU8, 1 > (m, *kp, grid, smem, st, pdl);
		case VDL2_FMT_CS8: return (int)launch_fmt < VDL2_FMT_CS8, 1 > (m, *kp, grid, smem, st, pdl);
		}
		return (int)cudaErrorInvalidValue;
	}
	switch (fmt) {
	case VDL2_FMT_CU8: return (int)launch_fmt < VDL2_FMT_CU8, 0 > (m, *kp, grid, smem, st, pdl);
	case VDL2_FMT_CS8: return (int)launch_fmt < VDL2_FMT_CS8, 0 > (m, *kp, grid, smem, st, pdl);
	case VDL2_FMT_CF32: return (int)launch_fmt < VDL2_FMT_CF32, 0 > (m, *kp, grid, smem, st, pdl);
	case VDL2_FMT_CS16: return (int)launch_fmt < VDL2_FMT_CS16, 0 > (m, *kp, grid, smem, st, pdl);
	case VDL2_FMT_F32REAL: return (int)launch_fmt < VDL2_FMT_F32REAL, 0 > (m, *kp, grid, smem, st, pdl);
	}
	return (int)cudaErrorInvalidValue;
}

template < int FMT, int DP > static cudaError_t occ_fmt(int smem, int *n)
{
	cudaFuncSetAttribute(vdl2::vdl2_frontend_kernel < FMT, DP, true >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	int a = 0, b = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, vdl2::vdl2_frontend_kernel < FMT, DP, true >, 32, smem);
	if (e != cudaSuccess)
		return e;
	cudaFuncSetAttribute(vdl2::vdl2_frontend_kernel < FMT, DP, false >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, vdl2::vdl2_frontend_kernel < FMT, DP, false >, 32, smem);
	*n = a < b ? a : b;	/* the per-warp scratch is sized from this: one grid size for both instantiations */
	return e;
}

extern "C" int vdl2_kernel_occupancy(int fmt, int dp4a, int smem, int *ctas_per_sm)
{
	if (dp4a == 2) {
		switch (fmt) {
		case VDL2_FMT_CU8: return (int)occ_fmt < VDL2_FMT_CU8, 2 > (smem, ctas_per_sm);
		case VDL2_FMT_CS8: return (int)occ_fmt < VDL2_FMT_CS8, 2 > (smem, ctas_per_sm);
		}
		return (int)cudaErrorInvalidValue;
	}
	if (dp4a) {
		switch (fmt) {
		case VDL2_FMT_CU8: return (int)occ_fmt < VDL2_FMT_CU8, 1 > (smem, ctas_per_sm);
		case VDL2_FMT_CS8: return (int)occ_fmt < VDL2_FMT_CS8, 1 > (smem, ctas_per_sm);
		}
		return (int)cudaErrorInvalidValue;
	}
	switch (fmt) {
	case VDL2_FMT_CU8: return (int)occ_fmt < VDL2_FMT_CU8, 0 > (smem, ctas_per_sm);
	case VDL2_FMT_CS8: return (int)occ_fmt < VDL2_FMT_CS8, 0 > (smem, ctas_per_sm);
	case VDL2_FMT_CF32: return (int)occ_fmt < VDL2_FMT_CF32, 0 > (smem, ctas_per_sm);
	case VDL2_FMT_CS16: return (int)occ_fmt < VDL2_FMT_CS16, 0 > (smem, ctas_per_sm);
	case VDL2_FMT_F32REAL: return (int)occ_fmt < VDL2_FMT_F32REAL, 0 > (smem, ctas_per_sm);
	}
	return (int)cudaErrorInvalidValue;
}

extern "C" int vdl2_channelise_launch(int fmt, const void *tmap, int nstreams, int cps, int nrows, int nbox, int sched_slot, const void *bt, const void *dt,
				      void *out, size_t out_pitch, unsigned *ticket, int grid, void *stream)
{
	vdl2::Vdl2ChanlParams cp;
	cp.nstreams = nstreams;
	cp.cps = cps;
	cp.nrows = nrows;
	cp.ntiles = (nrows + VDL2_ROWS_PER_TILE - 1) / VDL2_ROWS_PER_TILE;
	cp.nbox = nbox;
	cp.sched_slot = sched_slot;
	cp.bt = (const uint4 *)bt;
	cp.dt = (const int4 *)dt;
	cp.out = (float2 *) out;
	cp.out_pitch = out_pitch;
	cp.ticket = ticket;
	const CUtensorMap & m = *reinterpret_cast < const CUtensorMap * >(tmap);
	const int smem = MM_NST * MM_STAGE + 64;
	if (fmt == VDL2_FMT_CU8)
		vdl2::vdl2_channelise_kernel < VDL2_FMT_CU8 > <<<grid, 32, smem, (cudaStream_t) stream >>> (m, cp);
	else if (fmt == VDL2_FMT_CS8)
		vdl2::vdl2_channelise_kernel < VDL2_FMT_CS8 > <<<grid, 32, smem, (cudaStream_t) stream >>> (m, cp);
	else
		return (int)cudaErrorInvalidValue;
	return (int)cudaGetLastError();
}

extern "C" int vdl2_kernel_nsmid(unsigned *d_scratch_word, unsigned *out)
{
	vdl2::vdl2_nsmid_kernel <<< 1, 1 >>> (d_scratch_word);
	cudaError_t e = cudaMemcpy(out, d_scratch_word, 4, cudaMemcpyDeviceToHost);
	return (int)e;
}

extern "C" int vdl2_kernel_upload_tables(const Vdl2Tables * t)
{
	/* everything except the schedule slots (those are owned by vdl2_kernel_upload_sched) */
	cudaError_t e = cudaMemcpyToSymbol(c_tab, t, offsetof(Vdl2Tables, sched_slots));
	if (e != cudaSuccess)
		return (int)e;
	return (int)cudaMemcpyToSymbol(c_tab, t->hcol, sizeof t->hcol, offsetof(Vdl2Tables, hcol));
}

extern "C" int vdl2_kernel_upload_sched(int slot, const unsigned *sched)
{
	return (int)cudaMemcpyToSymbol(c_tab, sched, sizeof(unsigned) * VDL2_DUMPS_PER_ROW,
				       offsetof(Vdl2Tables, sched_slots) + sizeof(unsigned) * VDL2_DUMPS_PER_ROW * slot);
}
