/* Issue-rate micro-benchmark for the int8 tensor-core mixer (round 2): cycles per warp-instruction per SMSP for
   mma.sync.m16n8k32 u8*s8 (SASS IMMA.16832), ldmatrix.x4 (LDSM) and their mixes with the mixer's ALU work,
   4 one-warp CTAs per SMSP like the front-end kernel.  Also checks the fragment layout used by mix_rows_mma
   (vdl2_kernel.cu) against a scalar evaluation. */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t(&a)[4], uint32_t b0, uint32_t b1)
{
	asm volatile ("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};":"+r" (c[0]), "+r"(c[1]),
		      "+r"(c[2]), "+r"(c[3]):"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldsm4(uint32_t(&a)[4], uint32_t addr)
{
	asm volatile ("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];":"=r" (a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]):"r"(addr));
}

/* MODE 0: IMMA only, 4 independent accumulator sets; 1: LDSM only; 2: 4 LDSM + 4 IMMA (the mixer's per-dump tensor work);
   3: mode 2 + 40 ALU/FMA instructions (the estimated per-dump overhead) */
template < int MODE > __global__ void __launch_bounds__(32, 16) k(int *out, int iters, uint32_t seed)
{
	__shared__ __align__(1024) unsigned char sm[8192];
	for (int i = threadIdx.x; i < 2048; i += 32)
		reinterpret_cast < uint32_t * >(sm)[i] = i * 2654435761u ^ seed;
	__syncwarp();
	const int lane = threadIdx.x;
	const uint32_t base = (uint32_t) __cvta_generic_to_shared(sm);
	const uint32_t rowoff = ((lane >> 3) & 1) * 512 + (lane & 7) * 64, sw = (lane >> 1) & 3, cwl = lane >> 4;
	int c[4][4];
	uint32_t a[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) {
			c[i][j] = 0;
			a[i][j] = seed * (i + 3) + j + lane;
		}
	uint32_t b0 = seed ^ lane, b1 = seed + lane * 7;
	float f = 1.0f;
	uint32_t u = lane;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++) {
			if (MODE == 1 || MODE >= 2) {
				const uint32_t q0 = ((it + r) & 3) + cwl, q1 = q0 + 2;
				const uint32_t st = ((it + r) & 1) * 2048;
				const uint32_t ad0 = base + st + rowoff + (((q0 & 3) ^ sw) << 4), ad1 = base + st + rowoff + (((q1 & 3) ^ sw) << 4);
				ldsm4(a[0], ad0);
				ldsm4(a[1], ad0 + 1024);
				ldsm4(a[2], ad1);
				ldsm4(a[3], ad1 + 1024);
			}
			if (MODE == 0 || MODE >= 2) {
				imma(c[0], a[0], b0, b1);
				imma(c[1], a[1], b0, b1);
				imma(c[2], a[2], b1, b0);
				imma(c[3], a[3], b1, b0);
			}
			if (MODE == 3) {
#pragma unroll
				for (int q = 0; q < 20; q++) {
					f = fmaf(f, 1.0001f, 0.5f);
					u = (u ^ 0x9e3779b9u) + (u >> 3);
				}
			}
			if (MODE == 1) {
#pragma unroll
				for (int i = 0; i < 4; i++)
					c[i][0] += a[i][0] ^ a[i][1] ^ a[i][2] ^ a[i][3];
			}
		}
	}
	int s = __float_as_int(f) + u;
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++)
			s += c[i][j];
	out[blockIdx.x * 32 + threadIdx.x] = s;
}

template < int MODE > static void run(const char *name, int *d_out, double inst_per_iter)
{
	const int grid = 148 * 16, iters = 2000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k < MODE > <<<grid, 32 >>> (d_out, 10, 12345u);
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	k < MODE > <<<grid, 32 >>> (d_out, iters, 12345u);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	int clk = 0;
	cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	const double cyc = ms * 1e-3 * clk * 1e3;
	const double winst_per_smsp = 4.0 * iters * inst_per_iter;
	printf("%-40s %8.3f ms  %.2f cycles per counted warp-instruction per SMSP (%s)\n", name, ms, cyc / winst_per_smsp, cudaGetErrorString(cudaGetLastError()));
}

/* layout check: one warp, A = 16 x 32 bytes (u8), B = 32 x 8 (s8) */
__global__ void layout_check(const uint8_t * A, const int8_t * B, int *C)
{
	__shared__ __align__(1024) uint8_t sa[16 * 32];
	const int lane = threadIdx.x;
	for (int i = lane; i < 16 * 32; i += 32)
		sa[i] = A[i];
	__syncwarp();
	/* ldmatrix: matrix i = lane >> 3: rows 8 (i & 1) .., 16-byte column chunk i >> 1; row pitch 32 bytes here */
	uint32_t a[4];
	const uint32_t addr = (uint32_t) __cvta_generic_to_shared(sa) + (((lane >> 3) & 1) * 8 + (lane & 7)) * 32 + (lane >> 4) * 16;
	ldsm4(a, addr);
	const int g = lane >> 2, t = lane & 3;
	uint32_t b0 = 0, b1 = 0;
	for (int i = 0; i < 4; i++) {
		b0 |= (uint32_t) (uint8_t) B[(4 * t + i) * 8 + g] << (8 * i);
		b1 |= (uint32_t) (uint8_t) B[(16 + 4 * t + i) * 8 + g] << (8 * i);
	}
	int c[4] = { 1000, 2000, 1000, 2000 };
	imma(c, a, b0, b1);
	C[g * 8 + 2 * t] = c[0] - 1000;
	C[g * 8 + 2 * t + 1] = c[1] - 2000;
	C[(g + 8) * 8 + 2 * t] = c[2] - 1000;
	C[(g + 8) * 8 + 2 * t + 1] = c[3] - 2000;
}

int main()
{
	int *d_out;
	cudaMalloc(&d_out, sizeof(int) * 148 * 16 * 32);
	uint8_t hA[512];
	int8_t hB[256];
	int hC[128], ref[128];
	srand(7);
	for (int i = 0; i < 512; i++)
		hA[i] = rand() & 255;
	for (int i = 0; i < 256; i++)
		hB[i] = (int8_t) (rand() & 255);
	for (int r = 0; r < 16; r++)
		for (int n = 0; n < 8; n++) {
			int s = 0;
			for (int kk = 0; kk < 32; kk++)
				s += (int)hA[r * 32 + kk] * (int)hB[kk * 8 + n];
			ref[r * 8 + n] = s;
		}
	uint8_t *dA;
	int8_t *dB;
	int *dC;
	cudaMalloc(&dA, 512);
	cudaMalloc(&dB, 256);
	cudaMalloc(&dC, 512);
	cudaMemcpy(dA, hA, 512, cudaMemcpyHostToDevice);
	cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
	layout_check <<< 1, 32 >>> (dA, dB, dC);
	cudaMemcpy(hC, dC, 512, cudaMemcpyDeviceToHost);
	int bad = 0;
	for (int i = 0; i < 128; i++)
		bad += hC[i] != ref[i];
	printf("fragment layout check (ldmatrix.x4 + mma.m16n8k32.u8.s8, C-init): %s (%d mismatches) %s\n", bad ? "FAIL" : "OK", bad,
	       cudaGetErrorString(cudaGetLastError()));
	run < 0 > ("IMMA.16832.U8.S8 x4 indep", d_out, 32);
	run < 1 > ("LDSM.x4 (+1 ALU)", d_out, 32);
	run < 2 > ("4 LDSM + 4 IMMA, counted per IMMA", d_out, 32);
	run < 3 > ("4 LDSM + 4 IMMA + 40 ALU/FMA, per IMMA", d_out, 32);
	return bad ? 1 : 0;
}
