/* Micro-benchmark of the per-dump mixer body (24 cu8 IQ samples per lane, per-sample FFMA2 with the
   LO_HI swap) with four ways of fetching the warp-uniform oscillator values:
   0 LDCU (uniform index, constant bank)   1 LDC (vector index, constant bank)
   2 LDS.64 broadcast                      3 LDS.128 broadcast (two samples)
   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_ubench mix_ubench.cu */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#ifndef OCC
#define OCC 16
#endif
__constant__ float2 cw[6144];
__constant__ unsigned csched[84];

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
__device__ __forceinline__ float2 cvt(uint32_t w, int hi)
{
	const uint32_t magic = 0x4B000000u;
	float2 x;
	x.x = __uint_as_float(__byte_perm(w, magic, hi ? 0x7442 : 0x7440));
	x.y = __uint_as_float(__byte_perm(w, magic, hi ? 0x7443 : 0x7441));
	return fadd2(x, make_float2(-8388735.f, -8388735.f));
}

template < int VAR, int RCP > __global__ void __launch_bounds__(32, OCC) k(const int *slots, float2 * out, int iters, int base_param)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const int lane = threadIdx.x;
	uint32_t *st = reinterpret_cast < uint32_t * >(smem);
	for (int i = lane; i < 1536 / 4; i += 32)
		st[i] = (i * 2654435761u) ^ (blockIdx.x * 40503u);
	float2 *wsm = reinterpret_cast < float2 * >(smem + 1536);
	for (int i = lane; i < 192; i += 32)
		wsm[i] = cw[i];
	__syncwarp();
	int base = (VAR == 1) ? slots[blockIdx.x & 31] : base_param;
	float2 acc = make_float2(0.f, 0.f);
	const uint4 *row = reinterpret_cast < const uint4 * >(smem + lane * 48);
	for (int it = 0; it < iters; it++) {
		int wrun = 0;
#pragma unroll 1
		for (int dk = 0; dk < 84; dk++) {
			int w0;
			if (VAR == 4) {
				w0 = base + wrun;
				wrun += (dk % 5 == 4) ? 23 : 24;
				if (wrun >= 80)
					wrun -= 80;
			} else {
				const unsigned sk = csched[dk];
				w0 = base + (int)(sk & 127u);
			}
			const uint4 v0 = row[0], v1 = row[1], v2 = row[2];
			const uint32_t d[12] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w };
			float2 P0 = cw[base + 104 + dk], Q0 = P0, P1 = make_float2(0.f, 0.f), Q1 = P1;
			float re0 = P0.x, im0 = P0.y, re1 = 0.f, im1 = 0.f;
#pragma unroll
			for (int p = 0; p < 12; p++) {
				float2 xa, xb;
				if (RCP == 2 || RCP == 4) {	/* scalar bias removal */
					const uint32_t magic = 0x4B000000u;
					xa.x = __uint_as_float(__byte_perm(d[p], magic, 0x7440)) - 8388735.f;
					xa.y = __uint_as_float(__byte_perm(d[p], magic, 0x7441)) - 8388735.f;
					xb.x = __uint_as_float(__byte_perm(d[p], magic, 0x7442)) - 8388735.f;
					xb.y = __uint_as_float(__byte_perm(d[p], magic, 0x7443)) - 8388735.f;
				} else {
					xa = cvt(d[p], 0);
					xb = cvt(d[p], 1);
				}
				float2 Wa, Wb;
				if (VAR <= 1 || VAR == 4) {
					Wa = cw[w0 + 2 * p];
					Wb = cw[w0 + 2 * p + 1];
				} else if (VAR == 2) {
					Wa = wsm[(w0 & 127) + 2 * p];
					Wb = wsm[(w0 & 127) + 2 * p + 1];
				} else {
					const float4 W = *reinterpret_cast < const float4 * >(wsm + ((w0 & 126) + 2 * p));
					Wa = make_float2(W.x, W.y);
					Wb = make_float2(W.z, W.w);
				}
				if (RCP == 1 || RCP == 4) {	/* packed MACs */
					P0 = ffma2(xa, Wa, P0);
					Q0 = ffma2(make_float2(xa.y, xa.x), Wa, Q0);
					P1 = ffma2(xb, Wb, P1);
					Q1 = ffma2(make_float2(xb.y, xb.x), Wb, Q1);
				} else {	/* scalar MACs */
					re0 = fmaf(xa.x, Wa.x, re0);
					re0 = fmaf(-xa.y, Wa.y, re0);
					im0 = fmaf(xa.x, Wa.y, im0);
					im0 = fmaf(xa.y, Wa.x, im0);
					re1 = fmaf(xb.x, Wb.x, re1);
					re1 = fmaf(-xb.y, Wb.y, re1);
					im1 = fmaf(xb.x, Wb.y, im1);
					im1 = fmaf(xb.y, Wb.x, im1);
				}
			}
			if (RCP == 2 || RCP == 3) {
				P0 = make_float2(re0 + re1, 0.f);
				Q0 = make_float2(im0 + im1, 0.f);
				P1 = Q1 = make_float2(0.f, 0.f);
			}
			const float2 P = fadd2(P0, P1), Q = fadd2(Q0, Q1);
			acc.x += P.x - P.y;
			acc.y += Q.x + Q.y;
			__syncwarp();
		}
	}
	out[blockIdx.x * 32 + lane] = acc;
}

template < int VAR, int RCP > static void run(const char *name, int *d_slots, float2 * d_out, int iters)
{
	const int grid = 148 * OCC, smem = 1536 + 192 * 8;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k < VAR, RCP > <<<grid, 32, smem >>> (d_slots, d_out, 2, 0);
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	k < VAR, RCP > <<<grid, 32, smem >>> (d_slots, d_out, iters, 0);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double dumps = (double)grid * iters * 84;	/* dump-warps */
	const double samples = dumps * 32 * 24;
	printf("%-28s %8.3f ms  %7.1f ns/dump-warp/SMSP-slot  %9.0f Msamples/s (mixer only)  err=%s\n", name, ms,
	       ms * 1e6 / (dumps / (148 * 4)), samples / ms / 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv)
{
	const int iters = argc > 1 ? atoi(argv[1]) : 200;
	float2 h[6144];
	for (int i = 0; i < 6144; i++)
		h[i] = make_float2((float)((i * 37) % 101) / 101.f, (float)((i * 53) % 103) / 103.f);
	cudaMemcpyToSymbol(cw, h, sizeof h);
	unsigned sc[84];
	int s = 0;
	for (int kx = 0; kx < 84; kx++) {
		sc[kx] = (unsigned)(s % 80);
		s += (kx % 5 == 4) ? 23 : 24;
	}
	cudaMemcpyToSymbol(csched, sc, sizeof sc);
	int hs[32];
	for (int i = 0; i < 32; i++)
		hs[i] = (i % 8) * 192;
	int *d_slots;
	float2 *d_out;
	cudaMalloc(&d_slots, sizeof hs);
	cudaMemcpy(d_slots, hs, sizeof hs, cudaMemcpyHostToDevice);
	cudaMalloc(&d_out, sizeof(float2) * 148 * 32 * 32);
	run < 3, 1 > ("LDS.128 W, R1 FADD2+FFMA2", d_slots, d_out, iters);
	run < 3, 2 > ("LDS.128 W, R2 FADD+FFMA", d_slots, d_out, iters);
	run < 3, 3 > ("LDS.128 W, R3 FADD2+FFMA", d_slots, d_out, iters);
	run < 3, 4 > ("LDS.128 W, R4 FADD+FFMA2", d_slots, d_out, iters);
	run < 0, 1 > ("const W,   R1 FADD2+FFMA2", d_slots, d_out, iters);
	run < 0, 3 > ("const W,   R3 FADD2+FFMA", d_slots, d_out, iters);
	run < 2, 3 > ("LDS.64 W,  R3 FADD2+FFMA", d_slots, d_out, iters);
	return 0;
}
