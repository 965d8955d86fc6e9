/*
 * vdl2_mma_tables.h -- host-side tables of the int8 tensor-core mixer (mix_rows_mma, vdl2_kernel.cu).
 * Plain host C++, no CUDA: included by vdl2_host.cu (the product) and by tests/emul/mma_mix_host.cpp, which
 * replays the kernel's lane-level data path on the CPU against the oracle (test infrastructure).
 *
 * What is computed (d8psk.c:366-381 with rtl.c:287-289 in front): per dump k of a 1 ms row
 *     D_k = (1/nf) sum_{n in k} x[n] w[n mod N],   x = (u8 - 127.37f) + i(u8 - 127.37f),  w = the reference's float NCO table
 * as ONE integer matrix product per dump over all 32 rows of a tile:
 *     A  = 32 rows x 64 window bytes (the raw interleaved I,Q bytes, unsigned for cu8 / signed for cs8),
 *          window = the four 16-byte chunks j0 .. j0+3 holding the dump (j0 = first sample / 8)
 *     B  = 64 x 8 signed bytes: columns 0..2 = digits 2,1,0 of the real-part weights (wr for I bytes, -wi for Q bytes),
 *          columns 4..6 = digits of the imaginary-part weights (wi for I, wr for Q), columns 3 and 7 duplicates
 *          that are multiplied by zero afterwards; weights = round(w * 8355711) in balanced base-256 digits,
 *          zero outside the dump (the kernel masks the per-phase table with the dump's first/last sample)
 *     C0 = 0x4B400000 - 128 sum(B column over the dump)  (cu8: undoes the +128 of unsigned bytes exactly, in integers;
 *          0x4B400000 = bits of 1.5 * 2^23, so the int32 accumulator IS the float 12582912 + sum when it comes out)
 * The int32 sums are exact; the three digits are combined in fp32, scaled by 1 / (8355711 nf) and corrected by
 * delta * sum(w) for the 0.63 LSB between 128 and the reference's 127.37f (linear, evaluated in double here).
 */
#ifndef VDL2_MMA_TABLES_H
#define VDL2_MMA_TABLES_H
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#define VDL2_MM_MAGIC 0x4B400000	/* float bits of 12582912 = 1.5 * 2^23 */
/* oscillator weights are quantised to 1 / 8355711: the largest scale whose range [-1, 1] still fits three balanced base-256
   digits (127 * 65536 + 127 * 256 + 127); quantisation error <= 6.0e-8, the size of one rounding of the reference's own
   float products (the IDP.4A mixer of round 1 uses 2^-22: 1.2e-7) */
#define VDL2_MM_WSCALE 8355711.0
#define VDL2_MM_W 0x1u			/* sched bit: the dump's last chunk lies in a box not waited for yet */
#define VDL2_MM_R 0x2u			/* sched bit: the current box is finished after this dump */

struct Vdl2MmaU4 { uint32_t x, y, z, w; };
struct Vdl2MmaI4 { int32_t x, y, z, w; };

/* the reference's oscillator table, d8psk.c:353-357: wf[n] = cexpf(-n * Fo' * I) with Fo' rounded to float */
static inline void vdl2_nco_table(int Fo, unsigned fs, int nco_n, float *wr, float *wi)
{
	const float Fp = (float)((float)Fo / (float)(fs) * 2.0 * M_PI);
	for (int n = 0; n < nco_n; n++) {
		const float a = (float)(-n) * Fp;
		wr[n] = cosf(a);
		wi[n] = sinf(a);
	}
}

/* dump k of a row: first sample and length (d8psk.c:374-381: clk += 21; dump when clk >= SDRCLK) */
static inline int vdl2_dump_bounds(int row_samples, int sdrclk, int *start, int *len, int cap)
{
	int clk = 0, s0 = 0, k = 0;
	for (int n = 0; n < row_samples; n++) {
		clk += 21;
		if (clk >= sdrclk) {
			clk %= sdrclk;
			if (k < cap) {
				start[k] = s0;
				len[k] = n + 1 - s0;
			}
			s0 = n + 1;
			k++;
		}
	}
	return k;
}

/* can this (rate, clock) use the tensor-core mixer?  all dumps 23 or 24 samples, NCO period a multiple of 8 samples */
static inline bool vdl2_mma_usable(int row_samples, int sdrclk, int nco_n, int ndumps, int max_phases)
{
	if (nco_n % 8 || nco_n / 8 > max_phases || row_samples % nco_n || ndumps > 128)
		return false;
	int start[128], len[128];
	if (vdl2_dump_bounds(row_samples, sdrclk, start, len, 128) != ndumps)
		return false;
	for (int k = 0; k < ndumps; k++)
		if (len[k] != 23 && len[k] != 24)
			return false;
	return true;
}

/* per dump: W | R | bits 4-5 j0 & 3 | bits 8-13 6 p (window phase p = j0 mod (N/8); 64 * 6 p = byte offset of the phase in
   the B table) | bits 16-22 16 o | bits 23-31 16 e;  o = first sample of the dump inside its 32-sample window, e = o + length.
   The fields sit where the kernel needs them with the fewest instructions.  Returns the number of 64-byte boxes per row. */
static inline int vdl2_mma_build_sched(int row_samples, int sdrclk, int nco_n, int ndumps, unsigned *sched)
{
	int start[128], len[128];
	vdl2_dump_bounds(row_samples, sdrclk, start, len, 128);
	int maxbox = 0;		/* box 0 is waited for before the loop */
	for (int k = 0; k < ndumps; k++) {
		const int j0 = start[k] >> 3, o = start[k] & 7, e = o + len[k];
		const int b0 = j0 >> 2, b1 = (j0 + 3) >> 2;
		unsigned w = ((unsigned)(j0 & 3) << 4) | ((unsigned)(6 * (j0 % (nco_n / 8))) << 8) | ((unsigned)(16 * o) << 16) | ((unsigned)(16 * e) << 23);
		if (b1 > maxbox) {
			w |= VDL2_MM_W;
			maxbox = b1;
		}
		if (k + 1 < ndumps && ((start[k + 1] >> 3) >> 2) > b0)
			w |= VDL2_MM_R;
		sched[k] = w;
	}
	return maxbox + 1;
}

static inline void vdl2_mma_digits(int W, int dg[3])
{				/* W = dg[2] * 65536 + dg[1] * 256 + dg[0], every digit in [-128, 127] */
	dg[0] = ((W + 128) & 255) - 128;
	W = (W - dg[0]) / 256;
	dg[1] = ((W + 128) & 255) - 128;
	dg[2] = (W - dg[1]) / 256;
}

/* One channel.  bt[(p * 6 + c) * 4 + t] = the four B-fragment registers of lane (column c, t = lane & 3) for window
   phase p: {samples 2t,2t+1 | 8+2t,.. | 16+2t,.. | 24+2t,..}, each register = (w_I, w_Q) of two samples.
   dt[k * 4 + t] = {C0 of column 2t, C0 of column 2t+1, bits of 1/(8355711 nf), bits of the offset correction (re for t < 2, im else)};
   dt has ndumps + 1 groups (the kernel prefetches one dump ahead), the last one a copy of the first. */
static inline void vdl2_mma_build_chan(const float *wr, const float *wi, int nco_n, int row_samples, int sdrclk, int ndumps, bool cu8,
				       Vdl2MmaU4 * bt, Vdl2MmaI4 * dt)
{
	std::vector < int >dr(3 * nco_n), di(3 * nco_n), dni(3 * nco_n);
	for (int n = 0; n < nco_n; n++) {
		const int qr = (int)lrint((double)wr[n] * VDL2_MM_WSCALE), qi = (int)lrint((double)wi[n] * VDL2_MM_WSCALE);
		vdl2_mma_digits(qr, &dr[3 * n]);
		vdl2_mma_digits(qi, &di[3 * n]);
		vdl2_mma_digits(-qi, &dni[3 * n]);
	}
	/* weight byte of column c (0..2 real part, 3..5 imaginary part; digit 2 - c % 3) for NCO index n, I (par 0) or Q (par 1) byte */
	auto wbyte =[&](int n, int c, int par)->int {
		const int d = 2 - c % 3;
		if (c < 3)
			return par ? dni[3 * n + d] : dr[3 * n + d];
		return par ? dr[3 * n + d] : di[3 * n + d];
	};
	const int nph = nco_n / 8;
	for (int p = 0; p < nph; p++)
		for (int c = 0; c < 6; c++)
			for (int t = 0; t < 4; t++) {
				uint32_t r[4];
				for (int q = 0; q < 4; q++) {	/* q = 2 s + i: samples 16 s + 8 i + 2t, +1 */
					const int jA = 8 * q + 2 * t;
					uint32_t v = 0;
					for (int b = 0; b < 4; b++) {
						const int n = (8 * p + jA + (b >> 1)) % nco_n;
						v |= (uint32_t) (wbyte(n, c, b & 1) & 255) << (8 * b);
					}
					r[q] = v;
				}
				bt[(p * 6 + c) * 4 + t] = Vdl2MmaU4 { r[0], r[1], r[2], r[3] };
			}
	int start[128], len[128];
	vdl2_dump_bounds(row_samples, sdrclk, start, len, 128);
	const double delta = cu8 ? 128.0 - (double)(float)127.37 : 0.0;	/* x = (u - 128) + delta, rtl.c:287-289 */
	static const int col6_of[8] = { 0, 1, 2, 2, 3, 4, 5, 5 };
	for (int k = 0; k < ndumps; k++) {
		long long S[6] = { 0, 0, 0, 0, 0, 0 };
		double swr = 0, swi = 0;
		for (int j = 0; j < len[k]; j++) {
			const int n = (start[k] + j) % nco_n;
			for (int c = 0; c < 6; c++)
				S[c] += wbyte(n, c, 0) + wbyte(n, c, 1);
			swr += (double)wr[n];
			swi += (double)wi[n];
		}
		const double s = 1.0 / (double)len[k];
		const float sf = (float)(s / VDL2_MM_WSCALE);
		const float cre = (float)(delta * (swr - swi) * s), cim = (float)(delta * (swr + swi) * s);
		for (int t = 0; t < 4; t++) {
			Vdl2MmaI4 v;
			v.x = VDL2_MM_MAGIC - (cu8 ? 128 * (int)S[col6_of[2 * t]] : 0);
			v.y = VDL2_MM_MAGIC - (cu8 ? 128 * (int)S[col6_of[2 * t + 1]] : 0);
			memcpy(&v.z, &sf, 4);
			memcpy(&v.w, t < 2 ? &cre : &cim, 4);
			dt[k * 4 + t] = v;
		}
	}
	for (int t = 0; t < 4; t++)
		dt[ndumps * 4 + t] = dt[t];
}
#endif
