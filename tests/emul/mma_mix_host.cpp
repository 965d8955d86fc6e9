/*
 * mma_mix_host.cpp -- TEST INFRASTRUCTURE: lane-level CPU replay of the int8 tensor-core mixer (mix_rows_mma,
 * vdlm2dec_b200/csrc/vdl2_kernel.cu) on the product's own host tables (vdl2_mma_tables.h).
 *
 * It models exactly what the kernel relies on and nothing else: the 64B-swizzled TMA box layout with zero fill outside
 * the tensor, the stage ring with its W/R schedule bits, ldmatrix.x4 row addressing, the m16n8k32 fragment layouts
 * (checked on a B200 by tools/ubench/imma_ubench.cu), the B masks, the accumulator start values, the FFMA2 epilogue and the
 * exchange over a quad of dumps that leaves every lane with (re, im) of ONE dump of four rows.  tests/test_host_logic.py compares its dumps with the oracle's T1 tap, so the index
 * arithmetic and the tables are pinned in the CPU tier; the GPU tier pins the kernel itself.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "vdl2_common.h"
#include "vdl2_mma_tables.h"

#define MM_NST 4
#define MM_STAGE 2048

namespace {
struct Emul {
	std::vector < uint8_t > smem;	/* stages */
	int stage_box[MM_NST];	/* which box a stage holds (the model's stand-in for the mbarrier: reading a stage that does not hold
				   the expected box, or refilling one that was never waited for, is an error) */
	bool stage_waited[MM_NST];
	const uint8_t *rows;
	int nrows, row_bytes, row_samples;
	int errors = 0, loads = 0, waits = 1;

	void tma_load(int st, int box) {
		/* box = 32 samples (64 bytes) x 32 rows at sample offset box * 32; chunk c of row r at r*64 + ((c ^ ((r>>1)&3)) << 4) */
		for (int r = 0; r < 32; r++)
			for (int c = 0; c < 4; c++)
				for (int b = 0; b < 16; b++) {
					const long off = (long)box * 64 + c * 16 + b;
					uint8_t v = 0;
					if (r < nrows && off < row_bytes)
						v = rows[(size_t) r * row_bytes + off];
					smem[st * MM_STAGE + r * 64 + ((c ^ ((r >> 1) & 3)) << 4) + b] = v;
				}
		stage_box[st] = box;
		stage_waited[st] = false;
		loads++;
	}
};

inline uint32_t shl_clamp(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v << n; }
inline uint32_t shr_clamp(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v >> n; }
inline float as_float(int32_t v) { float f; memcpy(&f, &v, 4); return f; }
}

/* rows: nrows (<= 32) rows of row_samples interleaved 8-bit IQ samples; out: [32][84] complex floats (row-major, zero rows
   beyond nrows are computed like the kernel does: from zero-filled boxes).  Returns the number of protocol errors. */
extern "C" int emul_mma_mix(const uint8_t * rows, int nrows, unsigned fs, unsigned sdrclk, int Fo, int cu8, float *out)
{
	const int row_samples = (int)(fs / 1000), nco_n = (int)(fs / 25000), ND = VDL2_DUMPS_PER_ROW;
	if (!vdl2_mma_usable(row_samples, (int)sdrclk, nco_n, ND, VDL2_MM_PHASES))
		return -1;
	std::vector < float >wr(nco_n), wi(nco_n);
	vdl2_nco_table(Fo, fs, nco_n, wr.data(), wi.data());
	std::vector < Vdl2MmaU4 > bt(VDL2_MM_BT_ENTRIES);
	std::vector < Vdl2MmaI4 > dt(VDL2_MM_DT_ENTRIES);
	unsigned sched[VDL2_DUMPS_PER_ROW];
	vdl2_mma_build_chan(wr.data(), wi.data(), nco_n, row_samples, (int)sdrclk, ND, cu8 != 0, bt.data(), dt.data());
	const int nbox = vdl2_mma_build_sched(row_samples, (int)sdrclk, nco_n, ND, sched);

	Emul E;
	E.smem.assign(MM_NST * MM_STAGE, 0xEE);
	E.rows = rows;
	E.nrows = nrows;
	E.row_samples = row_samples;
	E.row_bytes = row_samples * 2;
	for (int b = 0; b < MM_NST; b++)
		E.tma_load(b, b);
	int st = 0, box = 0;
	E.stage_waited[0] = true;	/* the wait in front of the loop */
	static float ppair[2][32][4], va[32][4], keep[32][4];	/* partial sums of a pair of dumps; completed components after round A */
	if (ND % 4)
		return -2;	/* whole quads of dumps per row */
	for (int dk = 0; dk < ND; dk++) {
		const unsigned sk = sched[dk];
		const int st1 = (st + 1 == MM_NST) ? 0 : st + 1;
		if (sk & VDL2_MM_W) {
			if (E.stage_waited[st1] || E.stage_box[st1] != box + 1)
				E.errors++;
			E.stage_waited[st1] = true;
			E.waits++;
		}
		uint32_t A[32][4][4];	/* [lane][ldsm index 0: m0 s0, 1: m1 s0, 2: m0 s1, 3: m1 s1][reg] */
		/* ldmatrix.x4: lane l supplies the address of row l & 7 of matrix l >> 3; lane l receives from matrix j the
		   4 bytes at row l >> 2, byte offset 4 (l & 3) */
		uint32_t addr[4][32];
		for (int lane = 0; lane < 32; lane++) {
			const uint32_t sw = (lane >> 1) & 3, cwl = lane >> 4;
			const uint32_t rowoff = ((lane >> 3) & 1) * 512 + (lane & 7) * 64;
			const uint32_t base0 = rowoff + st * MM_STAGE, base1 = rowoff + st1 * MM_STAGE;
			const uint32_t q0 = ((sk >> 4) & 3u) + cwl, q1 = q0 + 2;
			if ((q0 >= 4 || q1 >= 4) && (!E.stage_waited[st1] || E.stage_box[st1] != box + 1))
				E.errors++;
			if (!E.stage_waited[st] || E.stage_box[st] != box)
				E.errors++;
			const uint32_t ad0 = (q0 >= 4 ? base1 : base0) + (((q0 & 3) ^ sw) << 4);
			const uint32_t ad1 = (q1 >= 4 ? base1 : base0) + (((q1 & 3) ^ sw) << 4);
			addr[0][lane] = ad0;
			addr[1][lane] = ad0 + 1024;
			addr[2][lane] = ad1;
			addr[3][lane] = ad1 + 1024;
		}
		for (int x = 0; x < 4; x++)
			for (int lane = 0; lane < 32; lane++)
				for (int j = 0; j < 4; j++) {
					const uint32_t rowaddr = addr[x][8 * j + (lane >> 2)];
					uint32_t v;
					memcpy(&v, &E.smem[rowaddr + 4 * (lane & 3)], 4);
					A[lane][x][j] = v;
				}
		/* B fragments + masks, accumulators, MMAs */
		int32_t C[32][2][4];
		uint32_t Bf[32][4];
		for (int lane = 0; lane < 32; lane++) {
			const int g = lane >> 2, t = lane & 3;
			const int col6 = (g >> 2) * 3 + ((g & 3) < 2 ? (g & 3) : 2);
			Vdl2MmaU4 B = bt[((sk >> 8) & 63u) * 4u + col6 * 4 + t];
			const int o16 = (int)((sk >> 16) & 127u), e16 = (int)(sk >> 23);
			const int m0c = -32 * t, m2c = 288 + 32 * t, m3c = 416 + 32 * t;
			B.x &= shl_clamp(0xffffffffu, (uint32_t) (o16 + m0c > 0 ? o16 + m0c : 0));
			B.z &= shr_clamp(0xffffffffu, (uint32_t) (m2c - e16 > 0 ? m2c - e16 : 0));
			B.w &= shr_clamp(0xffffffffu, (uint32_t) (m3c - e16 > 0 ? m3c - e16 : 0));
			Bf[lane][0] = B.x;
			Bf[lane][1] = B.y;
			Bf[lane][2] = B.z;
			Bf[lane][3] = B.w;
			const Vdl2MmaI4 dc = dt[dk * 4 + t];
			for (int m = 0; m < 2; m++) {
				C[lane][m][0] = C[lane][m][2] = dc.x;
				C[lane][m][1] = C[lane][m][3] = dc.y;
			}
		}
		/* mma.m16n8k32: A(row, k): a0 row g k 4t..; a1 row g+8; a2 row g k 16+4t; a3 row g+8 k 16+4t.  B(k, n): b0 k 4t.. n g; b1 k 16+4t.. */
		for (int m = 0; m < 2; m++)
			for (int s = 0; s < 2; s++) {
				int Am[16][32], Bm[32][8];
				const int x = m + 2 * s;	/* ldsm index */
				for (int lane = 0; lane < 32; lane++) {
					const int g = lane >> 2, t = lane & 3;
					for (int b = 0; b < 4; b++) {
						const int sh = 8 * b;
						auto byteA =[&](uint32_t v)->int { const uint8_t u = (uint8_t) (v >> sh); return cu8 ? (int)u : (int)(int8_t) u; };
						Am[g][4 * t + b] = byteA(A[lane][x][0]);
						Am[g + 8][4 * t + b] = byteA(A[lane][x][1]);
						Am[g][16 + 4 * t + b] = byteA(A[lane][x][2]);
						Am[g + 8][16 + 4 * t + b] = byteA(A[lane][x][3]);
						Bm[4 * t + b][g] = (int)(int8_t) (uint8_t) (Bf[lane][2 * s] >> sh);
						Bm[16 + 4 * t + b][g] = (int)(int8_t) (uint8_t) (Bf[lane][2 * s + 1] >> sh);
					}
				}
				for (int lane = 0; lane < 32; lane++) {
					const int g = lane >> 2, t = lane & 3;
					for (int ci = 0; ci < 4; ci++) {
						const int row = g + 8 * (ci >> 1), col = 2 * t + (ci & 1);
						int32_t acc = C[lane][m][ci];
						for (int k = 0; k < 32; k++)
							acc += Am[row][k] * Bm[k][col];
						C[lane][m][ci] = acc;
					}
				}
			}
		/* epilogue */
		float p[32][4];
		for (int lane = 0; lane < 32; lane++) {
			const int t = lane & 3;
			const float scx = (t & 1) ? 1.f : 65536.f, scy = (t & 1) ? 0.f : 256.f;
			const float nx = -12582912.f * scx, ny = -12582912.f * scy;
			for (int q = 0; q < 4; q++) {
				const int m = q >> 1, h = q & 1;
				const float yx = fmaf(as_float(C[lane][m][2 * h]), scx, nx), yy = fmaf(as_float(C[lane][m][2 * h + 1]), scy, ny);
				p[lane][q] = yx + yy;
			}
		}
		/* Exchange over a quad of dumps (the kernel's MM_QUAD_STORE path): the four lanes of a quad (same g) end up with ONE dump each
		   -- lane t with dump 4q + t -- of the four rows g, g + 8, g + 16, g + 24, so that a store instruction touches 8 rows with one
		   full 32-byte sector per row instead of 32 rows with half a sector each.
		   round A, once per pair of dumps (lane ^ 1 holds the other digits of the same component): a lane keeps the dump of the pair
		   with its own parity and sends its partial sums of the other one;
		   round B, once per quad (lane ^ 2 holds the other component): t < 2 keeps the first pair's dump, t >= 2 the second pair's. */
		for (int lane = 0; lane < 32; lane++)
			for (int q = 0; q < 4; q++)
				ppair[dk & 1][lane][q] = p[lane][q];
		if (dk & 1) {
			for (int lane = 0; lane < 32; lane++) {
				const int t = lane & 3, odd = t & 1;
				const Vdl2MmaI4 dc = dt[(dk - 1 + odd) * 4 + t];	/* the kept dump's scale and offset correction */
				const float sf = as_float(dc.z), corr = as_float(dc.w);
				for (int j = 0; j < 4; j++) {
					const float x = ppair[odd][lane ^ 1][j];	/* the partner sends its partial of the dump it does not keep = my parity */
					va[lane][j] = fmaf(ppair[odd][lane][j] + x, sf, corr);
				}
			}
			if (!(dk & 2)) {
				memcpy(keep, va, sizeof keep);
			} else {
				for (int lane = 0; lane < 32; lane++) {
					const int g = lane >> 2, t = lane & 3, hi = t >> 1;
					for (int j = 0; j < 4; j++) {
						const float x = hi ? va[lane ^ 2][j] : keep[lane ^ 2][j];	/* the partner (other hi) sends hi' ? keep : va */
						const float mine = hi ? va[lane][j] : keep[lane][j];
						const int row = g + 8 * j, d = dk - 3 + t;
						out[((size_t) row * ND + d) * 2] = hi ? x : mine;
						out[((size_t) row * ND + d) * 2 + 1] = hi ? mine : x;
					}
				}
			}
		}
		if (sk & VDL2_MM_R) {
			if (box + MM_NST < nbox)
				E.tma_load(st, box + MM_NST);
			box++;
			st = st1;
		}
	}
	/* every box issued must have been waited for exactly once: the kernel's barrier phase bits rely on it */
	if (E.loads != nbox || E.waits != nbox)
		E.errors++;
	for (int s = 0; s < MM_NST; s++)
		if (!E.stage_waited[s])
			E.errors++;
	return E.errors;
}

/* the PRODUCT's oscillator table (vdl2_nco_table, the function vdl2_host.cu builds every mixer table from) */
extern "C" int emul_nco_table(int Fo, unsigned fs, float *out)
{
	const int n = (int)(fs / 25000);
	std::vector < float >wr(n), wi(n);
	vdl2_nco_table(Fo, fs, n, wr.data(), wi.data());
	for (int i = 0; i < n; i++) {
		out[2 * i] = wr[i];
		out[2 * i + 1] = wi[i];
	}
	return n;
}
