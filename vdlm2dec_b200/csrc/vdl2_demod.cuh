/*
 * vdl2_demod.cuh -- phase 2 of the front-end kernel: the D8PSK demodulator state machine,
 * one warp per channel, LANES OVER TIME.
 *
 * Input: the channel's decimated 84 ksps stream for one tile (<= 2688 dumps) in shared
 * memory, preceded by the 16 dumps of history.  The reference runs this as a strictly
 * sequential per-sample state machine (d8psk.c:232-333); here it is restated so that a
 * warp evaluates 32 consecutive idle-mode steps (or 32 consecutive symbols) at once:
 *
 *   idle (WSYNC, d8psk.c:241-313): a step happens every 2nd dump with a constant tap phase
 *     r = clk mod 4; step n is a pure function of the 17-dump window (-> phase P_n) and of
 *     the 17 phases P_{n-64}, P_{n-60}, .. P_n.  Each lane runs
 *     the complete 17-point line fit for its own step in the reference's operation order;
 *     the trigger `perr < 4 && err > perr` (d8psk.c:292) is evaluated with the neighbour
 *     lane's err through shuffles and the first trigger is found with ballot/ffs; steps
 *     after it are discarded (the reference would not have run them in idle mode).
 *   burst (GETHEAD/GETDATA/GETFEC, d8psk.c:314-332, :67-209): a symbol every 8th dump, tap
 *     phase constant; lanes slice 32 symbols at once, soft-demap and descramble them; the
 *     25-bit header goes through a lane-per-state restatement of the 32-state max-product
 *     trellis (viterbi.c:37-96, per channel instead of the reference's global arrays);
 *     payload bits are packed to bytes with ballots and scattered to data[row][col] through
 *     the closed form of the reference's column-major de-interleave (d8psk.c:117-206).
 *
 * The file is written against the small `vw::` warp layer so that tests/emul can compile
 * the SAME source for the host with a fibre-based 32-lane warp emulator and check it
 * against the oracle without a GPU (test infrastructure; the product builds only the
 * CUDA variant).
 */
#ifndef VDL2_DEMOD_CUH
#define VDL2_DEMOD_CUH
#include "vdl2_common.h"
#include "vdl2_tables.h"

#ifdef __CUDACC__
#define VQ __device__ __forceinline__
#define VQ_COLD __device__ __noinline__	/* rare paths: kept out of line so the hot loops stay I-cache resident */
#ifdef VDL2_OUTLINE_RARE	/* A/B: measured slower on the burst workload (state forced through local memory around the calls) */
#define VQ_RARE __device__ __noinline__
#else
#define VQ_RARE __device__ __forceinline__
#endif
#ifdef VDL2_PASSA_OUTLINE	/* A/B: 8 % slower on the whole kernel */
#define VQ_PASSA __device__ __noinline__
#else
#define VQ_PASSA __device__ __forceinline__
#endif
#define VDL2_CONST __constant__
namespace vw {
VQ int lane() { return threadIdx.x & 31; }
VQ void sync() { __syncwarp(); }
template <class T> VQ T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <class T> VQ T shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
VQ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
VQ unsigned atomic_inc(unsigned *p) { return atomicAdd(p, 1u); }
VQ void fence() { __threadfence(); }
VQ int ffs(unsigned m) { return __ffs(m); }
VQ int popc(unsigned m) { return __popc(m); }
VQ float fma(float a, float b, float c) { return fmaf(a, b, c); }
#ifdef VDL2_LIBM_ATAN2	/* A/B: CUDA's atan2f (59 instructions per call, 10 % of all instructions of the kernel in round 1) */
VQ float atan2(float y, float x) { return atan2f(y, x); }
#else
/* atan2f for the phase of a filtered sample (cargf, d8psk.c:229): octant reduction with one MUFU.RCP, atan(a) = a + a z Q(z)
   with Q a degree-7 minimax fit on z = a^2 in [0,1], quadrant fix-ups with two-word constants.  26 instructions.  Against the
   exact value on 4e6 random vectors: max 3.5e-7 rad, rms 8.2e-8 (glibc atan2f, what the reference calls: max 3.4e-7, rms 7.5e-8),
   i.e. the same error class as the libm it replaces; (0, 0) gives 0 like cargf(0).  tools/fuzz_gpu.py runs through it. */
VQ float atan2(float y, float x)
{
	const float ax = fabsf(x), ay = fabsf(y);
	const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
	float rc;
	asm("rcp.approx.f32 %0, %1;":"=f"(rc):"f"(mx));
	const float a = mn * rc, z = a * a;
	/* Estrin's scheme: four independent FMAs, then two combining levels, instead of an 8-deep Horner chain */
	const float z2 = z * z, z4 = z2 * z2;
	const float q0 = fmaf(0.1999039649963379f, z, -0.33332985639572144f), q1 = fmaf(0.1057392954826355f, z, -0.1418597400188446f);
	const float q2 = fmaf(0.04112180322408676f, z, -0.073667012155056f), q3 = fmaf(0.002622236730530858f, z, -0.01513250358402729f);
	const float p = fmaf(fmaf(q3, z2, q2), z4, fmaf(q1, z2, q0));
	float r = fmaf(a * z, p, a);
	if (ay > ax)
		r = (1.5707963705062866f - r) + -4.371138828673793e-08f;
	if (x < 0.f)
		r = (3.1415927410125732f - r) + -8.742277657347586e-08f;
	if (mx == 0.f)
		r = 0.f;
	return copysignf(r, y);
}
#endif
#ifdef VDL2_SLOW_SINCOS
VQ void sincos(float a, float &s, float &c) { sincosf(a, &s, &c); }
#else
/* only the correlation screen uses this (phases in [-pi, pi], 1e-6 is plenty): SFU sine / cosine */
VQ void sincos(float a, float &s, float &c) { s = __sinf(a); c = __cosf(a); }
#endif
VQ float rsqrt(float a) { return rsqrtf(a); }
VQ unsigned f2bits(float a) { return __float_as_uint(a); }
VQ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
VQ float fsub(float a, float b) { return __fsub_rn(a, b); }
VQ float fadd(float a, float b) { return __fadd_rn(a, b); }
VQ float fmul(float a, float b) { return __fmul_rn(a, b); }
template <class T> VQ T ldcg(const T * p) { return __ldcg(p); }
template <class T> VQ T ldg(const T * p) { return __ldg(p); }
/* A pointer that went through the argument list of an out-of-line function is a generic pointer: every access
   becomes a generic LD/ST.  Re-derive it from the dynamic shared array so that the compiler knows the address
   space again (LDS / STS, vector widths) -- measured 8 % of the whole kernel on idle_passA. */
template <class T> VQ T *as_shared(T * p)
{
	extern __shared__ __align__(1024) unsigned char vdl2_smem_base[];
	return reinterpret_cast < T * >(vdl2_smem_base + (__cvta_generic_to_shared(p) - __cvta_generic_to_shared(vdl2_smem_base)));
}
/* Blackwell packed fp32 (two independent IEEE operations per instruction) */
VQ float2 fma2(float2 a, float2 b, float2 c)
{
	unsigned long long ra = *reinterpret_cast < unsigned long long *>(&a), rb = *reinterpret_cast < unsigned long long *>(&b);
	unsigned long long rc = *reinterpret_cast < unsigned long long *>(&c), rd;
	asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(rd):"l"(ra), "l"(rb), "l"(rc));
	return *reinterpret_cast < float2 * >(&rd);
}
VQ float2 add2(float2 a, float2 b)
{
	unsigned long long ra = *reinterpret_cast < unsigned long long *>(&a), rb = *reinterpret_cast < unsigned long long *>(&b), rd;
	asm("add.rn.f32x2 %0, %1, %2;":"=l"(rd):"l"(ra), "l"(rb));
	return *reinterpret_cast < float2 * >(&rd);
}
}
#else
#include "vdl2_emul.h"		/* tests/emul: host definitions of VQ, VDL2_CONST and vw:: */
#endif

/* constant tables (struct Vdl2Tables, vdl2_common.h), filled by the host at create time.
   This header is included by exactly one translation unit per build, so it defines it. */
VDL2_CONST Vdl2Tables c_tab;

namespace vdl2 {

#define VDL2_PI_F 3.14159274101257324f	/* smallest float above M_PI: (double)x > M_PI  <=>  x >= VDL2_PI_F */
#define VDL2_2PI_HI 6.28318548202514648f
#define VDL2_2PI_LO (-1.74845553146951715e-07f)
#define VDL2_PI_D 3.14159265358979323846

/* parity-check columns of the (25,20) header code (viterbi.c:29-35), from the constant table */
VQ int hcol(int n)
{
	return c_tab.hcol[n];
}

VQ unsigned revbits(unsigned in, int n)
{				/* d8psk.c:39-52 */
	unsigned out = 0;
	for (int i = 0; i < n; i++) {
		out = (out << 1) | (in & 1);
		in >>= 1;
	}
	return out;
}

/* Interpolating low-pass + phase (filteredphase, d8psk.c:219-230) for the dump at tile index d:
   window sd[d .. d+16] (sd carries 16 dumps of history in front), oldest first, taps
   mflt[clk], mflt[clk+4], ... < 65.  General tap phase: taps beyond the table are zeros, so the
   loop always runs 17 taps (x*0 adds nothing), and the 17 L2 loads are issued back to back. */
VQ float filt_phase_any(const float2 * sd, int d, int clk)
{
#ifdef VDL2_OLD_FILT
	float sr0 = 0.f, si0 = 0.f;
	int jj = 0;
	for (int i = clk; i < VDL2_MFLTLEN; i += 4, jj++) {
		const float m = c_tab.mflt[i];
		const float2 x = vw::ldcg(sd + d + jj);
		sr0 = vw::fma(x.x, m, sr0);
		si0 = vw::fma(x.y, m, si0);
	}
	return vw::atan2(si0, sr0);
#endif
	float2 x[17];
#pragma unroll
	for (int j = 0; j < 17; j++)
		x[j] = vw::ldcg(sd + d + j);
	float sr = 0.f, si = 0.f;
#pragma unroll
	for (int j = 0; j < 17; j++) {
		const int i = clk + 4 * j;
		const float m = c_tab.mflt[i < 67 ? i : 67];	/* entries 63..67 are zero */
		sr = vw::fma(x[j].x, m, sr);
		si = vw::fma(x[j].y, m, si);
	}
	return vw::atan2(si, sr);
}

/* 17-point least-squares line through the unwrapped (phase - unique word) sequence
   (d8psk.c:259-289).  ph[4*l] is the phase of symbol l.  Returns residual and slope. */
VQ_COLD float2 sync_fit_nv(const float *ph_generic)
{
	const float *ph = vw::as_shared(ph_generic);
	float Pr[VDL2_NBPH];
	float kf = 0.f, Pv, M;
	M = Pv = Pr[0] = ph[0] - c_tab.sync[0];
#pragma unroll
	for (int l = 1; l < VDL2_NBPH; l++) {
		const float Pc = ph[4 * l] - c_tab.sync[l];
		const float Pd = Pc - Pv;
		Pv = Pc;
		/* the reference keeps Pu = (float)(Pu -+ 2*M_PI); kf counts the net turns and
		   Pu is rebuilt as the correctly rounded kf*2pi (differs from the sequentially
		   rounded value by <= 1 ulp, see DESIGN.md "numerics") */
		kf += (Pd >= VDL2_PI_F) ? -1.f : ((Pd <= -VDL2_PI_F) ? 1.f : 0.f);
		const float Pu = vw::fma(kf, VDL2_2PI_HI, kf * VDL2_2PI_LO);
		Pr[l] = Pc + Pu;
		M += Pr[l];
	}
	M = vw::fdiv(M, 17.f);
	float fr = 0.f;
#pragma unroll
	for (int l = 0; l < VDL2_NBPH; l++) {
		Pr[l] -= M;
		fr = vw::fma(Pr[l], (float)(l - 8), fr);
	}
	fr = vw::fdiv(fr, 408.f);
	float err = 0.f;
#pragma unroll
	for (int l = 0; l < VDL2_NBPH; l++) {
		const float e = vw::fma(-(float)(l - 8), fr, Pr[l]);
		err = vw::fma(e, e, err);
	}
	return make_float2(err, fr);
}

VQ void sync_fit(const float *ph, float &err_out, float &fr_out)
{
	const float2 r = sync_fit_nv(ph);
	err_out = r.x;
	fr_out = r.y;
}

/* burst geometry from the header (d8psk.c:94-95, :139-162, :188-197) */
struct BurstGeom {
	int nbrow, nlbyte;	/* data phase */
	int frow, flast;	/* FEC phase: rows, bytes of the short last row (0 = all rows full) */
	int nd, nf;		/* bytes in each phase */
	int nsym;		/* symbols from the trigger to the end of the burst */
};

VQ BurstGeom burst_geom(int nbrow, int nlbyte)
{
	BurstGeom g;
	g.nbrow = nbrow;
	g.nlbyte = nlbyte;
	g.nd = nlbyte ? nlbyte * nbrow + (249 - nlbyte) * (nbrow - 1) : 249 * nbrow;
	if (nlbyte <= 2) {
		g.frow = nbrow - 1;
		g.flast = 0;
	} else {
		g.frow = nbrow;
		g.flast = nlbyte <= 30 ? 2 : (nlbyte <= 67 ? 4 : 0);
	}
	g.nf = g.flast ? g.flast * g.frow + (6 - g.flast) * (g.frow - 1) : 6 * g.frow;
	g.nsym = (25 + 8 * (g.nd + g.nf) + 2) / 3;
	return g;
}

/* byte number B of the burst payload -> offset in data[8][255] */
VQ int byte_slot(const BurstGeom & g, int B)
{
	int r, c;
	if (B < g.nd) {
		const int full = g.nlbyte * g.nbrow;
		if (g.nlbyte == 0 || B < full) {
			c = B / g.nbrow;
			r = B - c * g.nbrow;
		} else {
			const int b2 = B - full, w = g.nbrow - 1;
			const int q = b2 / w;
			c = g.nlbyte + q;
			r = b2 - q * w;
		}
	} else {
		const int b1 = B - g.nd;
		const int full = g.flast * g.frow;
		if (g.flast == 0 || b1 < full) {
			c = b1 / g.frow;
			r = b1 - c * g.frow;
		} else {
			const int b2 = b1 - full, w = g.frow - 1;
			const int q = b2 / w;
			c = g.flast + q;
			r = b2 - q * w;
		}
		c += 249;
	}
	return r * 255 + c;
}

/* 25-bit header through the 32-state max-product trellis, lane = state (viterbi.c:37-96).
   The update order of the reference (ascending source state, '1' branch before '0') decides
   ties; it is reproduced by ordering the two candidates of a target state by source index. */
VQ_COLD unsigned header_decode(const float *hv)
{
	const int s = vw::lane();
	double pb = (s == 0) ? 1.0 : 0.0;
	unsigned hist = 0;
#pragma unroll 1
	for (int n = 0; n < 25; n++) {	/* once per burst: not unrolled, the kernel has to stay inside the instruction cache */
		const float V = hv[n];
		const int src1 = s ^ hcol(n);
		const double p1 = vw::shfl(pb, src1);
		const double c1 = p1 * (double)V;
		const double c0 = pb * (1.0 - (double)V);
		const bool v1 = (p1 != 0.0), v0 = (pb != 0.0);
		double nw = 0.0;
		unsigned bit = 0;
		if (src1 < s) {
			if (v1 && c1 > nw) { nw = c1; bit = 1; }
			if (v0 && c0 > nw) { nw = c0; bit = 0; }
		} else {
			if (v0 && c0 > nw) { nw = c0; bit = 0; }
			if (v1 && c1 > nw) { nw = c1; bit = 1; }
		}
		pb = nw;
		hist |= bit << n;
	}
	unsigned bits = 0, b = 1;
	int st = 0;
#pragma unroll 1
	for (int n = 25; n > 0; n--) {
		const unsigned h = vw::shfl(hist, st);
		const unsigned bit = (h >> (n - 1)) & 1u;
		if (bit) {
			bits |= b;
			st ^= hcol(n - 1);
		}
		b <<= 1;
	}
	return bits;
}

/* ---------------------------------------------------------------------------------------
 * demod_tile: consume dumps [0, nd) of one tile.
 *   sd   : float2[16 + nd] (L2-resident scratch of this warp), sd[16 + i] = dump i, sd[0..15] = history
 *   S    : shared scratch of the idle search; S.pht[0..63] = the 64 previous idle-mode phases (oldest
 *          first); on return nph phases were appended (the new history is S.pht[nph .. nph+63])
 *   hv   : shared, float[28]; header soft bits collected so far
 *   st   : the channel's state in HBM (read at entry, written back at exit by the caller
 *          through the ChanRegs copy below)
 * ------------------------------------------------------------------------------------- */
struct ChanRegs {		/* warp-uniform working copy of the scalar state */
	float perr, p2err, pfr, df, P1, ppm;
	int clk, state, symidx, nbrow, nlbyte, bytes_done, bitacc, nbitacc;
	long long sync_dump;
	unsigned n_steps, n_syncs, n_syms;
};

VQ_RARE void emit_block(const Vdl2KParams & kp, int ch, const ChanRegs & R, long long end_dump, int chn, int Fr)
{
	const int lane = vw::lane();
	unsigned slot = 0;
	if (lane == 0)
		slot = vw::atomic_inc(kp.outq_count);
	slot = vw::shfl(slot, 0);
	unsigned *cur = (unsigned *)(kp.curblk + (size_t) ch * 2048);
	if (slot < kp.outq_cap) {
		Vdl2BlockRec *rec = kp.outq + slot;
		unsigned *dst = (unsigned *)rec->data;	/* 2040 bytes = 510 words, 8-byte aligned */
		for (int i = lane; i < 510; i += 32)
			dst[i] = vw::ldcg(cur + i);
		if (lane == 0) {
			rec->sync_dump = R.sync_dump;
			rec->end_dump = end_dump;
			rec->chn = chn;
			rec->Fr = Fr;
			rec->ppm = R.ppm;
			rec->nbrow = R.nbrow;
			rec->nlbyte = R.nlbyte;
		}
	} else if (lane == 0) {
		vw::atomic_inc(kp.dropped);
	}
	vw::sync();
	for (int i = lane; i < 512; i += 32)	/* fresh zeroed block (calloc in vdlm2.c:201) */
		cur[i] = 0;
	vw::sync();
}

/* scratch of the idle-mode search (shared memory; the kernel lends it the idle TMA stages) */
#define VDL2_PHT_LEN (VDL2_PHHIST + VDL2_TILE_DUMPS / 2 + 32)
#define VDL2_CAND_CAP 192
#define VDL2_CAND0_CAP 96	/* candidates among the first VDL2_PRE_SKIP steps of a speculatively screened tile */
#define VDL2_PRE_SKIP 96	/* steps whose screen needs phases of the previous tile (64 + 8 rounded up to a batch) */
#define VDL2_WIN_LEN 80
struct IdleScratch {
	float *pht;		/* [VDL2_PHT_LEN]: pht[0..63] = the 64 phases before the tile, then one per idle step */
	float2 *vw;		/* [96]: differential phasors v_t = z_t conj(z_{t-4}), sliding window */
	unsigned short *cand;	/* [VDL2_CAND_CAP]: steps that passed the screen */
	unsigned short *cand0;	/* [VDL2_CAND0_CAP]: same for the head of a speculatively screened tile (idle_prepass) */
	float2 *win;		/* [VDL2_WIN_LEN]: the dumps one batch of 32 steps filters (copied from the L2-resident stream) */
	unsigned char *hb;	/* [VDL2_TILE_DUMPS / 8]: decisions of burst symbols taken ahead of the chain (BurstPre) */
};

/* filter only (no phase): 17 taps, steady-state tap phase */
VQ void filt17(const float2 * sd, int d, const float *m, float &sr, float &si)
{
#ifndef VDL2_FILT_DUAL	/* one accumulator, the order of the scalar loop (17 dependent FFMA2) */
	float2 s = make_float2(0.f, 0.f);
#pragma unroll
	for (int j = 0; j < 17; j++)
		s = vw::fma2(sd[d + j], make_float2(m[j], m[j]), s);
	sr = s.x;
	si = s.y;
#else
	/* A/B (-DVDL2_FILT_DUAL): two accumulators (even / odd taps), half the dependent chain.  Measured on B200: 2.682 vs 2.685 ms per
	   step, i.e. nothing -- so the order of the reference's loop stays the default */
	float2 a = make_float2(0.f, 0.f), b = a;
#pragma unroll
	for (int j = 0; j < 16; j += 2) {
		a = vw::fma2(sd[d + j], make_float2(m[j], m[j]), a);
		b = vw::fma2(sd[d + j + 1], make_float2(m[j + 1], m[j + 1]), b);
	}
	a = vw::fma2(sd[d + 16], make_float2(m[16], m[16]), a);
	const float2 s = vw::add2(a, b);
	sr = s.x;
	si = s.y;
#endif
}

/* three independent L2 loads per lane: window entries lane, lane+32, lane+64 starting at sd[g0] */
VQ void win_fetch(const float2 * sd, int g0, int glim, float2 & wa, float2 & wb, float2 & wc)
{
	const int l = vw::lane();
	const float2 z = make_float2(0.f, 0.f);
	const int ga = g0 + l, gb = g0 + l + 32, gc = g0 + l + 64;
	wa = ga < glim ? vw::ldcg(sd + ga) : z;
	wb = gb < glim ? vw::ldcg(sd + gb) : z;
	wc = (l < VDL2_WIN_LEN - 64 && gc < glim) ? vw::ldcg(sd + gc) : z;
}

VQ float2 cmul_conj(float2 a, float2 b)
{				/* a * conj(b) */
	return make_float2(vw::fma(a.x, b.x, a.y * b.y), vw::fma(a.y, b.x, -(a.x * b.y)));
}

/* ---------------------------------------------------------------------------------------
 * idle_run: all idle-mode steps from `pos` to the end of the tile (or to the first trigger).
 *
 * The reference evaluates the full 17-point fit at every step (d8psk.c:259-289) although the
 * trigger needs it only around residuals below 4.0.  Here every step still gets its exact
 * phase P (the Ph ring must stay exact), but the fit runs only where a cheap NECESSARY
 * condition for err < 4 holds:
 *   with u_l = exp(i(Pr_l - Pr_{l-1})) = z_l conj(z_{l-1}) exp(-i(SW_l - SW_{l-1})), z = exp(iP),
 *   err = sum e_l^2 < 4  =>  |sum_{l=1..16} u_l| >= 16 - (lambda_max/2) err > 8.07
 *   (lambda_max = 2 - 2cos(16pi/17) of the 17-node path Laplacian; DESIGN.md "idle screen").
 * SW_l - SW_{l-1} = (2 q_l + 1) pi/8 with q_l the unique word, so the sum is a 16-tap
 * correlation of the differential phasors with multiples of pi/4: adds and swaps only.
 * Steps passing |sum|^2 >= 62 (|sum| >= 7.87, margin for fp32) are "candidates"; the exact
 * fit runs lane-parallel over the candidate list; the first candidate whose exact err < 4
 * switches the rest of the run to the exact lane-per-step search, two steps early so that
 * p2err/perr/pfr are the reference's values at the trigger.  Everything before it provably
 * cannot trigger.  The last two steps of a run are always fitted (state for the next tile).
 * ------------------------------------------------------------------------------------- */
#define VDL2_SCREEN_THR2 62.0f

template < bool TAPS > VQ_RARE void idle_trigger(const Vdl2KParams & kp, int ch, int Fr, ChanRegs & R, const float2 * sd, long long dump_base, int dT, float eT,
		     float pT, float p2T, float fT)
{				/* d8psk.c:292-308 */
	R.state = VDL2_ST_GETHEAD;
	R.symidx = 0;
	R.df = fT;
	R.ppm = (float)((double)vw::fmul(10500.0f, R.df) / (2.0 * VDL2_PI_D * (double)Fr) * 1e6);
	const float num = vw::fmul(4.0f, vw::fadd(vw::fsub(p2T, vw::fmul(4.0f, pT)), vw::fmul(3.0f, eT)));
	const float den = vw::fadd(vw::fsub(p2T, vw::fmul(2.0f, pT)), eT);
	const float of = vw::fdiv(num, den);
	int nclk = (int)roundf(of);
	nclk = nclk < 0 ? 0 : (nclk > 64 ? 64 : nclk);	/* of is in [4,12] for finite input */
	R.clk = nclk;
	R.P1 = filt_phase_any(sd, dT, nclk);
	R.perr = R.p2err = 500.f;
	R.sync_dump = dump_base + dT;
	if ((TAPS && (kp.taps & VDL2_TAP_SYNCS_BIT))) {
		if (vw::lane() == 0 && R.n_syncs < kp.cap_syncs) {
			Vdl2SyncRec s;
			s.dump = R.sync_dump;
			s.clk = nclk;
			s.df = R.df;
			s.ppm = R.ppm;
			s.P1 = R.P1;
			kp.tap_syncs[(size_t) ch * kp.cap_syncs + R.n_syncs] = s;
		}
		R.n_syncs++;
	}
}

/* Pass A of the idle search for the steps of batches [b_begin, b_end): exact phase of every step into
   ph[VDL2_PHHIST + k]; if `screen`, the correlation screen marks candidates (steps k >= k_cand0 only) in
   cand[].  `have_hist`: ph[0..63] holds the 64 phases before the run and sd[] the 16 dumps before the tile;
   without it (idle_prepass) the first steps produce garbage that the caller never uses. */
VQ_PASSA int idle_passA(const float2 * sd, IdleScratch S, float *ph, int r, int p0, int N, int nd, int b_begin, int b_end, bool screen,
		       bool have_hist, int k_cand0, unsigned short *cand, int cand_cap)
{				/* out of line: three call sites, and the kernel has to stay inside the instruction cache.
				   Returns the number of candidates (> cand_cap: the list overflowed) */
	const int lane = vw::lane();
	S.pht = vw::as_shared(S.pht);
	S.vw = vw::as_shared(S.vw);
	S.win = vw::as_shared(S.win);
	ph = vw::as_shared(ph);
	cand = vw::as_shared(cand);
	int ncand = 0;
	float m[17];
#pragma unroll
	for (int j = 0; j < 17; j++)
		m[j] = c_tab.mflt[r + 4 * j];	/* r + 64 <= 67: zero padded */
	float2 zlast = make_float2(1.f, 0.f);
	if (screen && have_hist) {
		float2 zA, zB;
		vw::sincos(ph[b_begin + lane], zA.y, zA.x);
		vw::sincos(ph[b_begin + 32 + lane], zB.y, zB.x);
		const float2 a4 = make_float2(vw::shfl_up(zA.x, 4), vw::shfl_up(zA.y, 4));
		const float2 b4u = make_float2(vw::shfl_up(zB.x, 4), vw::shfl_up(zB.y, 4));
		const float2 b4w = make_float2(vw::shfl(zA.x, (lane + 28) & 31), vw::shfl(zA.y, (lane + 28) & 31));
		S.vw[lane] = cmul_conj(zA, a4);	/* entries 0..3 are never read */
		S.vw[32 + lane] = cmul_conj(zB, lane >= 4 ? b4u : b4w);
		zlast = zB;
		vw::sync();
	}
	float2 wa, wb, wc;
	win_fetch(sd, p0 + 2 * b_begin, VDL2_HIST + nd, wa, wb, wc);
	for (int b0 = b_begin; b0 < b_end; b0 += 32) {
		const int k = b0 + lane;
		/* the 79 dumps this batch filters, sd[p0 + 2*b0 .. +78], were fetched from L2 one batch ahead */
		S.win[lane] = wa;
		S.win[lane + 32] = wb;
		if (lane < VDL2_WIN_LEN - 64)
			S.win[lane + 64] = wc;
		vw::sync();
		if (b0 + 32 < b_end)
			win_fetch(sd, p0 + 2 * (b0 + 32), VDL2_HIST + nd, wa, wb, wc);
		float sr = 1.f, si = 0.f;
		if (k < N)
			filt17(S.win, 2 * lane, m, sr, si);
		const float P = vw::atan2(si, sr);
		if (k < N)
			ph[VDL2_PHHIST + k] = P;
		if (screen) {
			const float mag2 = vw::fma(sr, sr, si * si);
			const bool ok = (mag2 > 1e-30f) && (mag2 < 1e30f);
			const float rn = vw::rsqrt(ok ? mag2 : 1.f);
			const float2 z = make_float2(sr * rn, si * rn);
			const float2 z4u = make_float2(vw::shfl_up(z.x, 4), vw::shfl_up(z.y, 4));
			const float2 z4w = make_float2(vw::shfl(zlast.x, (lane + 28) & 31), vw::shfl(zlast.y, (lane + 28) & 31));
			S.vw[64 + lane] = cmul_conj(z, lane >= 4 ? z4u : z4w);
			zlast = z;
			vw::sync();
			/* correlation with the unique word: q = 0,3,2,4,0,1,6,4,1,7,2,5,6,5,7,3 (multiples of pi/4) */
			float2 A0 = make_float2(0.f, 0.f), A2 = A0, B0 = A0, B2 = A0;
#define VDL2_ACC(L, Q) { const float2 w = S.vw[lane + 4 * (L)]; \
	if ((Q) == 0) A0 = vw::add2(A0, w); else if ((Q) == 4) A0 = vw::fma2(w, mone, A0); \
	else if ((Q) == 2) A2 = vw::add2(A2, w); else if ((Q) == 6) A2 = vw::fma2(w, mone, A2); \
	else if ((Q) == 1) B0 = vw::add2(B0, w); else if ((Q) == 5) B0 = vw::fma2(w, mone, B0); \
	else if ((Q) == 3) B2 = vw::add2(B2, w); else B2 = vw::fma2(w, mone, B2); }
			const float2 mone = make_float2(-1.f, -1.f);
			VDL2_ACC(1, 0) VDL2_ACC(2, 3) VDL2_ACC(3, 2) VDL2_ACC(4, 4) VDL2_ACC(5, 0) VDL2_ACC(6, 1) VDL2_ACC(7, 6) VDL2_ACC(8, 4)
			VDL2_ACC(9, 1) VDL2_ACC(10, 7) VDL2_ACC(11, 2) VDL2_ACC(12, 5) VDL2_ACC(13, 6) VDL2_ACC(14, 5) VDL2_ACC(15, 7) VDL2_ACC(16, 3)
#undef VDL2_ACC
			/* C = A0 - i A2 + exp(-i pi/4) (B0 - i B2) */
			const float cx = A0.x + A2.y, cy = A0.y - A2.x;
			const float dx = B0.x + B2.y, dy = B0.y - B2.x;
			const float sx = cx + 0.70710678f * (dx + dy), sy = cy + 0.70710678f * (dy - dx);
			const float c2 = vw::fma(sx, sx, sy * sy);
			const bool cnd = (k < N) && (k >= k_cand0) && (!ok || !(c2 < VDL2_SCREEN_THR2) || k >= N - 2);
			const unsigned cm = vw::ballot(cnd);
			const int slot = ncand + vw::popc(cm & ((1u << lane) - 1u));
			if (cnd && slot < cand_cap)
				cand[slot] = (unsigned short)k;
			ncand += vw::popc(cm);
			/* slide the window by 32 steps */
			const float2 w1 = S.vw[32 + lane], w2 = S.vw[64 + lane];
			vw::sync();
			S.vw[lane] = w1;
			S.vw[32 + lane] = w2;
		}
		vw::sync();
	}
	return ncand;
}

/* Speculative pass A (see the kernel: it runs while the warp would otherwise wait for the previous tile of
   its channel): demod_tile is entered once with `spec` set and a GUESS of the channel's tick clock at the
   start of the tile in R.clk; idle_run then only runs pass A without history (phases of all but the first
   steps, screen of all but the first VDL2_PRE_SKIP) and records what it assumed.  The real call verifies the
   guess and falls back to the normal path otherwise.  One code path serves both so that pass A, the largest
   loop of phase 2, exists once in the kernel (out of line it ran 8 % slower, twice inline it does not fit
   the instruction cache). */
struct IdlePre {
	int pos0;		/* tile dump the speculative run starts at: 0, or the dump after the last symbol of a burst that ends inside the tile */
	int valid, clk, ncand;
	bool overflow;
	int used;		/* set by idle_run when the speculative results were accepted (statistics) */
};

/* One symbol: differential phase -> soft-table index -> soft bits (d8psk.c:209-216).  Returns both hard decisions every bit can
   end in, because the descrambler (d8psk.c:54-65) either keeps the soft bit or replaces it by 1 - v before the comparison with
   0.5: bit q = (v_q > 0.5), bit 3 + q = (1 - v_q > 0.5) (not complements of each other at v = 0.5, which the table contains). */
VQ unsigned sym_decide(const Vdl2KParams & kp, float Pn, float Pp, float df, float &D, int &gi, float *v)
{
	D = vw::fsub(vw::fsub(Pn, Pp), df);
	if (D >= VDL2_PI_F)
		D = (float)((double)D - 2.0 * VDL2_PI_D);
	if (D <= -VDL2_PI_F)
		D = (float)((double)D + 2.0 * VDL2_PI_D);
	gi = (int)roundf((float)(128.0 * (double)D / VDL2_PI_D + 128.0));	/* d8psk.c:213 */
	gi = gi < 0 ? 0 : (gi > 256 ? 256 : gi);
	unsigned ab = 0;
#pragma unroll
	for (int q = 0; q < 3; q++) {
		/* per-lane index: from global memory (L1/L2), not the constant bank, which would serialise the 32
		   distinct addresses of a symbol batch -- the largest stall of the burst path on the per-channel chain */
		v[q] = kp.soft ? vw::ldg(kp.soft + q * 260 + gi) : c_tab.soft[q][gi];
		ab |= (v[q] > 0.5f ? 1u : 0u) << q;
		ab |= (vw::fsub(1.0f, v[q]) > 0.5f ? 1u : 0u) << (3 + q);
	}
	return ab;
}

/* Burst phases ahead of the chain.  Inside a burst nothing but the phase of the previous symbol is carried from one symbol
   to the next (d8psk.c:314-332): the phase of the symbol at tile dump d is a function of the tile's own dumps and the tap
   phase r alone as soon as its 17-dump window lies inside the tile (d >= 16).  Once a header is decoded the channel's forecast
   says where the burst ends and with which r, so a tile that starts inside it computes those phases (the filter, its 17
   strided L2 loads and the atan2: most of a symbol batch) while it waits for the previous tile; the burst loop then takes
   them from bph[d >> 3] whenever the ACTUAL symbol grid and tap phase are the ones assumed -- never otherwise, so a stale or
   torn forecast costs time, not correctness.  The phases live at the top of S.pht, filled downwards (VDL2_BPH): the phases of
   the idle steps after the end of a burst at dump d -- the speculative pass A of the rest of the tile runs ahead of the chain
   too -- end below index 1408 - d / 2, the burst's above 1439 - d / 8. */
struct BurstPre {
	int valid, r, d0, dlast;	/* phases of the symbols at tile dumps d0, d0 + 8, ... <= dlast are in bph[d >> 3] */
	float df;		/* and, for all but the first of them, the decisions (sym_decide) under this frequency offset in hb[d >> 3] */
};
#define VDL2_BPH(d) (VDL2_PHT_LEN - 1 - ((d) >> 3))
#define VDL2_BPRE_NS 28		/* symbols per staging pass of burst_prephase: 8 * 28 + 9 = 233 dumps, swizzled into < 240 of the 248 float2 of scratch */
#define VDL2_BPRE_BUF ((96 * 8 + VDL2_WIN_LEN * 8 + (VDL2_CAND_CAP + VDL2_CAND0_CAP) * 2) / 8)	/* float2 from IdleScratch.vw to the end of cand0 */
static_assert(((8 * VDL2_BPRE_NS + 9 - 1) | 15) < VDL2_BPRE_BUF, "staged burst windows must fit vw .. cand0");

VQ_RARE void burst_prephase(const Vdl2KParams & kp, const float2 * sd, const IdleScratch & S, BurstPre & bp, int d0, int dlast, int r, float df)
{
	float *pht = vw::as_shared(S.pht);
	unsigned char *hb = vw::as_shared(S.hb);
	const int dmin = (d0 & 7) + 16;
#ifdef VDL2_BPRE_DIRECT	/* A/B: every lane fetches its own 17-dump window from L2 -- symbols are 8 dumps = 64 bytes apart, so each of the 17
				   loads of a batch touches 16 lines: 7 % of all LSU wavefronts of the kernel on the bench workload (ncu, round 2 v19) */
#pragma unroll 1
	for (int d = dmin + 8 * vw::lane(); d <= dlast; d += 256)
		pht[VDL2_BPH(d)] = filt_phase_any(sd, d, r);
#else
	/* The windows of VDL2_BPRE_NS consecutive symbols are one contiguous run of 8 NS + 9 dumps: staged with coalesced loads in the
	   idle-search scratch (vw, win, cand, cand0: contiguous, unused until the idle pass that follows), then every lane filters its
	   symbol from shared memory.  Entry i sits at i ^ ((i >> 4) & 15), which spreads the 64-byte symbol stride over all banks.
	   Same taps, same order of the sums as filt_phase_any: the phases are bit identical. */
	{
		float2 *buf = vw::as_shared(S.vw);
		const int lane = vw::lane();
		float m[17];
#pragma unroll
		for (int j = 0; j < 17; j++)
			m[j] = c_tab.mflt[r + 4 * j < 67 ? r + 4 * j : 67];	/* entries 63..67 are zero */
#pragma unroll 1
		for (int db = dmin; db <= dlast; db += 8 * VDL2_BPRE_NS) {
			const int ns = ((dlast - db) >> 3) + 1 < VDL2_BPRE_NS ? ((dlast - db) >> 3) + 1 : VDL2_BPRE_NS;
			const int len = 8 * ns + 9;
			vw::sync();
			for (int i = lane; i < len; i += 32)
				buf[i ^ ((i >> 4) & 15)] = vw::ldcg(sd + db + i);
			vw::sync();
			if (lane < ns) {
				float sr = 0.f, si = 0.f;
#pragma unroll
				for (int j = 0; j < 17; j++) {
					const int i = 8 * lane + j;
					const float2 x = buf[i ^ ((i >> 4) & 15)];
					sr = vw::fma(x.x, m[j], sr);
					si = vw::fma(x.y, m[j], si);
				}
				pht[VDL2_BPH(db + 8 * lane)] = vw::atan2(si, sr);
			}
		}
	}
#endif
	vw::sync();
#pragma unroll 1
	for (int d = dmin + 8 + 8 * vw::lane(); d <= dlast; d += 256) {
		float D, v[3];
		int gi;
		hb[d >> 3] = (unsigned char)sym_decide(kp, pht[VDL2_BPH(d)], pht[VDL2_BPH(d - 8)], df, D, gi, v);
	}
	bp.valid = dlast >= dmin;
	bp.r = r;
	bp.d0 = dmin;
	bp.dlast = dlast;
	bp.df = df;
	vw::sync();
}

template < bool TAPS > VQ void idle_run(const Vdl2KParams & kp, int ch, int Fr, ChanRegs & R, const float2 * sd, const IdleScratch & S, int nd,
		 long long dump_base, int &pos, int &nph, IdlePre & pre, bool spec)
{
	const int lane = vw::lane();
	if (R.clk >= 8)
		R.clk &= 7;	/* unreachable for finite input (clk < 8 whenever the burst clock was sane) */
	const int c4 = R.clk >= 4;
	const int p0 = pos + (c4 ? 0 : 1);	/* dump of the first step: a step every 2nd dump (d8psk.c:248-250) */
	const int r = c4 ? R.clk - 4 : R.clk;	/* tap phase of every step of the run */
	const bool use_pre = !spec && pre.valid && pos == pre.pos0 && nph == 0 && pre.clk == R.clk && R.perr >= 4.0f
	    && !(TAPS && (kp.taps & VDL2_TAP_STEPS_BIT)) && !(TAPS && (kp.flags & VDL2_FLAG_NO_SCREEN));
	pre.valid = 0;		/* a prepass covers the first run of the tile only */
	if (use_pre)
		pre.used = 1;
	if (p0 >= nd) {
		R.clk += 4 * (nd - pos);
		pos = nd;
		return;
	}
	const int N = ((nd - 1 - p0) >> 1) + 1;	/* steps in this run */
	float *ph = S.pht + nph;	/* ph[0..63] = history, ph[64 + k] = phase of step k */

	const bool exact_only = (TAPS && (kp.taps & VDL2_TAP_STEPS_BIT)) || (TAPS && (kp.flags & VDL2_FLAG_NO_SCREEN)) || !(R.perr >= 4.0f) || N < 8;

	/* ---- pass A: exact phase of every step; screen ---- */
	int ncand = 0, ncand0 = 0;
	bool overflow = false;
	if (spec && N < VDL2_PRE_SKIP + 32) {
		pos = nd;	/* too short to be worth it: no speculation */
		return;
	}
#ifdef VDL2_CHAIN_STATS	/* debug build: the idle step with an accepted speculation in two parts (head again with the real history | exact fits) */
	unsigned long long cs_a0 = 0, cs_a1 = 0;
	if (use_pre)
		asm volatile ("mov.u64 %0, %%globaltimer;":"=l" (cs_a0));
#endif
	{
		/* one call site, three uses:
		   spec     all batches, no history: candidates only from step VDL2_PRE_SKIP on;
		   use_pre  the head of the tile again, now that the previous tile's dumps and phases are known;
		   else     the plain full pass */
		const int b_end = use_pre ? VDL2_PRE_SKIP : N;
		const int nc = idle_passA(sd, S, ph, r, p0, N, nd, 0, b_end, spec || use_pre || !exact_only, !spec, spec ? VDL2_PRE_SKIP : 0,
					  use_pre ? S.cand0 : S.cand, use_pre ? VDL2_CAND0_CAP : VDL2_CAND_CAP);
		if (spec) {
			pre.ncand = nc;
			pre.overflow = nc > VDL2_CAND_CAP;
			pre.clk = R.clk;
			pre.valid = 1;
			pos = nd;
			return;
		}
		if (use_pre) {
			ncand0 = nc;
			ncand = pre.ncand;
			overflow = pre.overflow || nc > VDL2_CAND0_CAP;
#ifdef VDL2_CHAIN_STATS
			asm volatile ("mov.u64 %0, %%globaltimer;":"=l" (cs_a1));
#endif
		} else {
			ncand = nc;
			overflow = nc > VDL2_CAND_CAP;
		}
	}

	/* ---- pass B: exact fit of the candidates, in time order, until the first err < 4 ---- */
	int s0 = exact_only || overflow ? 0 : -1;	/* first step of the exact search; -1: no trigger possible in this run */
	float eN1 = 0.f, eN2 = 0.f, fN1 = 0.f;
	if (s0 < 0) {
		const int ntot = ncand0 + ncand;
		for (int c0 = 0; c0 < ntot; c0 += 32) {
			const int ci = c0 + lane;
			const bool have = ci < ntot;
			const int k = have ? (int)(ci < ncand0 ? S.cand0[ci] : S.cand[ci - ncand0]) : 0;
			float err, fr;
			sync_fit(ph + k, err, fr);
			const unsigned hm = vw::ballot(have && err < 4.0f);
			const unsigned l1 = vw::ballot(have && k == N - 1), l2 = vw::ballot(have && k == N - 2);
			if (l1) {
				eN1 = vw::shfl(err, vw::ffs(l1) - 1);
				fN1 = vw::shfl(fr, vw::ffs(l1) - 1);
			}
			if (l2)
				eN2 = vw::shfl(err, vw::ffs(l2) - 1);
			if (hm) {
				const int first = vw::shfl(k, vw::ffs(hm) - 1);
				s0 = first >= 2 ? first - 2 : 0;
				break;
			}
		}
	}
	if (s0 < 0) {		/* nothing below 4.0 anywhere: the whole run is committed */
#ifdef VDL2_CHAIN_STATS
		if (use_pre && lane == 0) {
			unsigned long long cs_a2;
			asm volatile ("mov.u64 %0, %%globaltimer;":"=l" (cs_a2));
			atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 54), cs_a1 - cs_a0);
			atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 56), cs_a2 - cs_a1);
			atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 58), (unsigned long long)(ncand0 + ncand));
			atomicAdd(reinterpret_cast < unsigned long long *>(kp.ticket + 60), 1ull);
		}
#endif
		nph += N;
		R.p2err = eN2;
		R.perr = eN1;
		R.pfr = fN1;
		pos = p0 + 2 * (N - 1) + 1;
		R.clk = r;
		return;
	}

	/* ---- exact search from step s0: one lane per step, reference trigger logic ---- */
	if (s0 > 0)
		R.perr = R.p2err = 500.f;	/* steps s0-1, s0-2 had err >= 4: any such value gives the same decisions */
	for (int b0 = s0; b0 < N; b0 += 32) {
		const int nb = (N - b0) < 32 ? (N - b0) : 32;
		const int k = b0 + lane;
		float err, fr;
		sync_fit(ph + (k < N ? k : 0), err, fr);
		const float e1 = vw::shfl_up(err, 1), e2 = vw::shfl_up(err, 2), f1 = vw::shfl_up(fr, 1);
		const float perr_l = lane >= 1 ? e1 : R.perr;
		const float p2err_l = lane >= 2 ? e2 : (lane == 1 ? R.perr : R.p2err);
		const float pfr_l = lane >= 1 ? f1 : R.pfr;
		const bool trig = (lane < nb) && (perr_l < 4.0f) && (err > perr_l);
		const unsigned tm = vw::ballot(trig);
		const int K = tm ? vw::ffs(tm) : nb;	/* steps committed, trigger step included */
		if ((TAPS && (kp.taps & VDL2_TAP_STEPS_BIT)) && lane < K) {
			const unsigned idx = R.n_steps + lane;
			if (idx < kp.cap_steps) {
				Vdl2StepRec s;
				s.dump = dump_base + p0 + 2 * k;
				s.P = ph[VDL2_PHHIST + k];
				const bool tl = tm && lane == K - 1;
				s.err = tl ? -1.0f : err;
				s.fr = tl ? pfr_l : fr;
				s.pad = 0;
				kp.tap_steps[(size_t) ch * kp.cap_steps + idx] = s;
			}
		}
		R.n_steps += (TAPS && (kp.taps & VDL2_TAP_STEPS_BIT)) ? K : 0;
		if (!tm) {
			const float eL = vw::shfl(err, K - 1), fL = vw::shfl(fr, K - 1);
			const float eP = vw::shfl(err, K >= 2 ? K - 2 : 0);
			R.p2err = K >= 2 ? eP : R.perr;
			R.perr = eL;
			R.pfr = fL;
		} else {
			const int T = K - 1;
			const float eT = vw::shfl(err, T), pT = vw::shfl(perr_l, T);
			const float p2T = vw::shfl(p2err_l, T), fT = vw::shfl(pfr_l, T);
			const int dT = p0 + 2 * (b0 + T);
			idle_trigger < TAPS > (kp, ch, Fr, R, sd, dump_base, dT, eT, pT, p2T, fT);
			nph += b0 + T + 1;	/* the Ph ring stops at the trigger step (d8psk.c:254-255) */
			pos = dT + 1;
			return;
		}
	}
	nph += N;
	pos = p0 + 2 * (N - 1) + 1;
	R.clk = r;
}

template < bool TAPS > VQ void demod_tile(const Vdl2KParams & kp, int ch, int chn, int Fr, ChanRegs & R, float2 * sd, const IdleScratch & S, float *hv,
		   int nd, long long dump_base, int &nph, IdlePre & pre, bool spec, BurstPre & bp)
{
	const int lane = vw::lane();
	int pos = spec ? pre.pos0 : 0;
	nph = 0;
	unsigned char *curblk = kp.curblk + (size_t) ch * 2048;

	while (pos < nd) {
		if (R.state == VDL2_ST_WSYNC) {
			if (!spec)
				bp.valid = 0;	/* idle steps write S.pht: whatever burst phases were computed ahead are gone */
			idle_run < TAPS > (kp, ch, Fr, R, sd, S, nd, dump_base, pos, nph, pre, spec);
		} else {
			/* ---- burst: up to 32 symbols at dumps ds0, ds0+8, ... (d8psk.c:314-332) ---- */
			const int c = R.clk;
			const int k0 = (c >= 28) ? 1 : ((35 - c) >> 2);
			const int r = c + 4 * k0 - 32;
			const int ds0 = pos + k0 - 1;
			if (ds0 >= nd) {
				R.clk += 4 * (nd - pos);
				pos = nd;
				break;
			}
			const bool head = (R.state == VDL2_ST_GETHEAD);
			BurstGeom g;
			if (!head)
				g = burst_geom(R.nbrow, R.nlbyte);
			int nb = ((nd - 1 - ds0) >> 3) + 1;
			const int remain = head ? 9 - R.symidx : g.nsym - R.symidx;
			nb = nb < remain ? nb : remain;
			nb = nb < 32 ? nb : 32;
			if (r >= 4)
				nb = 1;	/* irregular clock (only after a non-finite timing estimate) */
			const bool grid_ok = bp.valid && r == bp.r && ((ds0 - bp.d0) & 7) == 0;
			const bool dec_ok = grid_ok && !head && bp.dlast >= bp.d0 + 8 && vw::f2bits(R.df) == vw::f2bits(bp.df)
			    && !(TAPS && (kp.taps & VDL2_TAP_SYMS_BIT));
			if (dec_ok && ds0 >= bp.d0 + 8 && ds0 <= bp.dlast) {
				/* ---- every symbol from here to the end of the tile (or of the burst) was decided ahead of the chain (BurstPre):
				   what is left is the descrambler and the bytes, one pass with a lane per BYTE of the bit stream
				   (the R.nbitacc bits left over, then 3 bits per symbol) instead of a batch per 32 symbols ---- */
				int n = ((bp.dlast - ds0) >> 3) + 1;
				n = n < remain ? n : remain;
				const int nbits = R.nbitacc + 3 * n;
				const int total = g.nd + g.nf;
				int nbytes = nbits >> 3;
				if (R.bytes_done + nbytes > total)
					nbytes = total - R.bytes_done;
				const unsigned char *hb = vw::as_shared(S.hb) + (ds0 >> 3);
				unsigned last = 0;	/* byte number nbytes of the stream: the bits left over */
#pragma unroll 1
				for (int b0 = 0; b0 <= nbytes; b0 += 32) {
					const int b = b0 + lane;
					const int sp = 8 * b - R.nbitacc;	/* symbol bit the byte starts at; it spans at most 4 symbols */
					const int j0 = sp > 0 ? (sp * 43691) >> 17 : 0;	/* sp / 3 */
					unsigned h4 = 0;
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const int j = j0 + k;
						if (j < n && b <= nbytes) {
							const unsigned ab = hb[j];
							const int bb = 3 * (R.symidx + j);	/* descrambler bits of the symbol, d8psk.c:54-65 */
							const unsigned w0 = kp.soft ? vw::ldg(kp.scr + ((bb >> 5) & (VDL2_SCR_WORDS - 1))) : c_tab.scr[(bb >> 5) & (VDL2_SCR_WORDS - 1)];
							const unsigned w1 = kp.soft ? vw::ldg(kp.scr + (((bb >> 5) + 1) & (VDL2_SCR_WORDS - 1))) : c_tab.scr[((bb >> 5) + 1) & (VDL2_SCR_WORDS - 1)];
							const unsigned s3 = (unsigned)((((unsigned long long)w1 << 32) | w0) >> (bb & 31)) & 7u;
							h4 |= ((ab & 7u & ~s3) | ((ab >> 3) & s3)) << (3 * k);
						}
					}
					const unsigned byte = (sp >= 0 ? (h4 >> (sp - 3 * j0)) : ((unsigned)R.bitacc | (h4 << (-sp)))) & 0xffu;
					if (b < nbytes)
						curblk[byte_slot(g, R.bytes_done + b)] = (unsigned char)byte;
					const unsigned lm = vw::ballot(b == nbytes);
					if (lm)
						last = vw::shfl(byte, vw::ffs(lm) - 1);
				}
				int left = nbits - 8 * nbytes;
				if (left > 7)
					left = 0;	/* only at the end of the burst: discarded (d8psk.c:203) */
				const int dl = ds0 + 8 * (n - 1);
				R.bitacc = (int)(last & ((1u << left) - 1u));
				R.nbitacc = left;
				R.bytes_done += nbytes;
				R.symidx += n;
				R.P1 = vw::as_shared(S.pht)[VDL2_BPH(dl)];
				R.clk = r;
				pos = dl + 1;
				if (R.bytes_done >= total) {
					vw::fence();
					vw::sync();
					emit_block(kp, ch, R, dump_base + dl, chn, Fr);	/* decodeVdlm2 hand-off, d8psk.c:201 */
					R.state = VDL2_ST_WSYNC;
				}
				continue;
			}
			if (dec_ok && ds0 < bp.d0 + 8) {	/* the symbols in front of the decided ones first, so that the pass above takes all the rest */
				const int nfirst = (bp.d0 + 8 - ds0) >> 3;
				nb = nb < nfirst ? nb : nfirst;
			}
			const int d = ds0 + 8 * lane;
			float Pn = 0.f;
			{
				const bool have = grid_ok && d >= bp.d0 && d <= bp.dlast;	/* phase computed ahead of the chain (BurstPre) */
				if (have)
					Pn = vw::as_shared(S.pht)[VDL2_BPH(d)];
				else if (lane < nb)
					Pn = filt_phase_any(sd, d, r);
				if (lane >= nb)
					Pn = 0.f;
			}
			float Pp = vw::shfl_up(Pn, 1);
			if (lane == 0)
				Pp = R.P1;
			const int si = R.symidx + lane;
			float D = 0.f, v[3] = { 0.f, 0.f, 0.f };
			int gi = 128;
			const unsigned ab = sym_decide(kp, Pn, Pp, R.df, D, gi, v);
			/* descrambler, d8psk.c:54-65: bits 3 si .. 3 si + 2 of the sequence */
			const int b0 = 3 * si;
			const unsigned w0 = kp.soft ? vw::ldg(kp.scr + ((b0 >> 5) & (VDL2_SCR_WORDS - 1))) : c_tab.scr[(b0 >> 5) & (VDL2_SCR_WORDS - 1)];
			const unsigned w1 = kp.soft ? vw::ldg(kp.scr + (((b0 >> 5) + 1) & (VDL2_SCR_WORDS - 1))) : c_tab.scr[((b0 >> 5) + 1) & (VDL2_SCR_WORDS - 1)];
			const unsigned s3 = (unsigned)((((unsigned long long)w1 << 32) | w0) >> (b0 & 31)) & 7u;
			float V[3];
			unsigned hard = 0;
#pragma unroll
			for (int q = 0; q < 3; q++) {
				const unsigned sb = (s3 >> q) & 1u;
				V[q] = sb ? vw::fsub(1.0f, v[q]) : v[q];
				hard |= ((ab >> (sb ? 3 + q : q)) & 1u) << q;
			}
			const float Plast = vw::shfl(Pn, nb - 1);

			if ((TAPS && (kp.taps & VDL2_TAP_SYMS_BIT)) && lane < nb) {
				const unsigned idx = R.n_syms + lane;
				if (idx < kp.cap_syms) {
					Vdl2SymRec s;
					s.dump = dump_base + d;
					s.D = D;
					s.P = Pn;
					s.gi = gi;
					s.v[0] = v[0];
					s.v[1] = v[1];
					s.v[2] = v[2];
					s.state_after = 0;
					s.pad = 0;
					kp.tap_syms[(size_t) ch * kp.cap_syms + idx] = s;
				}
			}
			R.n_syms += (TAPS && (kp.taps & VDL2_TAP_SYMS_BIT)) ? nb : 0;

			const int dl = ds0 + 8 * (nb - 1);	/* dump of the last symbol of the batch */
			if (head) {
				/* header bits 0..26 (d8psk.c:77-116): first three forced to 0 */
				if (lane < nb) {
#pragma unroll
					for (int q = 0; q < 3; q++) {
						const int b = 3 * si + q;
						hv[b] = (b < 3) ? 0.f : V[q];
					}
				}
				vw::sync();
				R.symidx += nb;
				R.P1 = Plast;
				R.clk = r;
				pos = dl + 1;
				if (R.symidx == 9) {
					const unsigned w = header_decode(hv) >> 5;
					const unsigned len = revbits(w, 17);
					R.nbrow = (int)(len / 1992u) + 1;
					R.nlbyte = (int)((len % 1992u + 7u) / 8u);
					long long idle_from = dump_base + dl + 1;	/* rejected header: idle again right away */
					if (len < 96u || R.nbrow > 8) {
						R.state = VDL2_ST_WSYNC;	/* d8psk.c:97-107 */
					} else {
						idle_from += 8LL * (burst_geom(R.nbrow, R.nlbyte).nsym - 9);	/* one symbol every 8 dumps */
						R.state = VDL2_ST_GETDATA;
						R.bytes_done = 0;
						/* bits 25 and 26 are already payload (d8psk.c:117-123) */
						R.bitacc = (hv[25] > 0.5f ? 1 : 0) | (hv[26] > 0.5f ? 2 : 0);
						R.nbitacc = 2;
					}
					if (kp.state && lane == 0) {	/* forecast for the tiles after this burst (Vdl2ChanState.fc_*) */
						kp.state[ch].fc_dump = idle_from;
						kp.state[ch].fc_clk = r;
						kp.state[ch].fc_df = R.df;
					}
				}
			} else {
				/* payload: pack 3 bits per lane into bytes, LSB first (d8psk.c:117-206) */
				/* the bit stream of the batch = the R.nbitacc bits left over + 3 bits per lane; byte number `lane` of it starts at
				   symbol bit sp and takes its 8 bits from at most 4 consecutive symbols */
				const int nbits = R.nbitacc + 3 * nb;
				const int total = g.nd + g.nf;
				int nbytes = nbits >> 3;
				if (R.bytes_done + nbytes > total)
					nbytes = total - R.bytes_done;
				const int sp = 8 * lane - R.nbitacc;
				const int j0 = sp > 0 ? (sp * 43691) >> 17 : 0;	/* sp / 3 */
				const unsigned h4 = vw::shfl(hard, j0 & 31) | (vw::shfl(hard, (j0 + 1) & 31) << 3) | (vw::shfl(hard, (j0 + 2) & 31) << 6)
				    | (vw::shfl(hard, (j0 + 3) & 31) << 9);
				const unsigned byte = (sp >= 0 ? (h4 >> (sp - 3 * j0)) : ((unsigned)R.bitacc | (h4 << (-sp)))) & 0xffu;
				if (lane < nbytes)
					curblk[byte_slot(g, R.bytes_done + lane)] = (unsigned char)byte;
				/* bits left over for the next batch: the head of the byte that would come next */
				int left = nbits - 8 * nbytes;
				if (left > 7)
					left = 0;	/* only at the end of the burst: discarded (d8psk.c:203) */
				const unsigned acc = vw::shfl(byte, nbytes & 31) & ((1u << left) - 1u);
				R.bitacc = (int)acc;
				R.nbitacc = left;
				R.bytes_done += nbytes;
				R.symidx += nb;
				R.P1 = Plast;
				R.clk = r;
				pos = dl + 1;
				if (R.bytes_done >= total) {
					vw::fence();
					vw::sync();
					emit_block(kp, ch, R, dump_base + dl, chn, Fr);	/* decodeVdlm2 hand-off, d8psk.c:201 */
					R.state = VDL2_ST_WSYNC;
				}
			}
		}
	}
}

}				/* namespace vdl2 */
#endif
