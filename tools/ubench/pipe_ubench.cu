/* Issue-rate micro-benchmark: cycles per warp-instruction per SMSP for the instructions the mixer uses
   (FFMA2, FADD2, FFMA, PRMT and their mixes), 4 warps per SMSP, 8 independent chains per warp. */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
	unsigned long long ra = *reinterpret_cast < unsigned long long *>(&a), rb = *reinterpret_cast < unsigned long long *>(&b);
	unsigned long long rc = *reinterpret_cast < unsigned long long *>(&c), rd;
	asm volatile ("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(rd):"l"(ra), "l"(rb), "l"(rc));
	return *reinterpret_cast < float2 * >(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
	unsigned long long ra = *reinterpret_cast < unsigned long long *>(&a), rb = *reinterpret_cast < unsigned long long *>(&b), rd;
	asm volatile ("add.rn.f32x2 %0, %1, %2;":"=l"(rd):"l"(ra), "l"(rb));
	return *reinterpret_cast < float2 * >(&rd);
}
__device__ __forceinline__ float ffma(float a, float b, float c)
{
	float d;
	asm volatile ("fma.rn.f32 %0, %1, %2, %3;":"=f"(d):"f"(a), "f"(b), "f"(c));
	return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
	uint32_t d;
	asm volatile ("prmt.b32 %0, %1, %2, %3;":"=r"(d):"r"(a), "r"(b), "r"(s));
	return d;
}
__device__ __forceinline__ uint32_t dp4a_us(uint32_t a, uint32_t b, uint32_t c)
{
	int d;
	asm volatile ("dp4a.u32.s32 %0, %1, %2, %3;":"=r"(d):"r"(a), "r"(b), "r"(c));
	return (uint32_t) d;
}
__device__ __forceinline__ uint32_t dp2a_us(uint32_t a, uint32_t b, uint32_t c)
{
	int d;
	asm volatile ("dp2a.lo.u32.s32 %0, %1, %2, %3;":"=r"(d):"r"(a), "r"(b), "r"(c));
	return (uint32_t) d;
}
/* MODE: 0 FFMA2  1 FADD2  2 FFMA  3 PRMT  4 FFMA2+PRMT 1:1  5 FFMA2+FADD2+PRMT 2:1:2 (the mixer mix)  6 FFMA+PRMT 1:1
         7 FFMA2 with the same B operand (register reuse)  8 FFMA2 swap form */
template < int MODE > __global__ void __launch_bounds__(32, 16) k(float2 * out, int iters, float2 seed)
{
	float2 a[8], b = seed, x = make_float2(seed.y, seed.x);
	const unsigned su = __float_as_uint(seed.x) | 1u, sv = __float_as_uint(seed.y);
	uint32_t u[8];
#pragma unroll
	for (int i = 0; i < 8; i++) {
		a[i] = make_float2(threadIdx.x + i, i);
		u[i] = threadIdx.x * 77 + i;
	}
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				if (MODE == 0) a[i] = ffma2(a[i], b, x);
				if (MODE == 1) a[i] = fadd2(a[i], b);
				if (MODE == 2) a[i].x = ffma(a[i].x, b.x, x.x);
				if (MODE == 3) u[i] = prmt(u[i], 0x4B000000u, 0x7440 + (i & 3));
				if (MODE == 4) { a[i] = ffma2(a[i], b, x); u[i] = prmt(u[i], 0x4B000000u, 0x7440 + (i & 3)); }
				if (MODE == 5) { a[i] = ffma2(a[i], b, x); u[i] = prmt(u[i], 0x4B000000u, 0x7440 + (i & 3));
					if (i & 1) a[i] = fadd2(a[i], b); }
				if (MODE == 6) { a[i].x = ffma(a[i].x, b.x, x.x); u[i] = prmt(u[i], 0x4B000000u, 0x7440 + (i & 3)); }
				if (MODE == 7) a[i] = ffma2(x, b, a[i]);
				if (MODE == 8) a[i] = ffma2(make_float2(a[i].y, a[i].x), b, x);
				if (MODE == 9) u[i] = __float_as_uint((float)((u[i] >> 8) & 0xffu));
				if (MODE == 10) { u[i] = __float_as_uint((float)((u[i] >> 8) & 0xffu)); a[i].x = ffma(a[i].x, b.x, x.x); a[i].y = ffma(a[i].y, b.x, x.x); }
				if (MODE == 11) { u[i] = __float_as_uint((float)((u[i] >> 8) & 0xffu)); a[i] = ffma2(a[i], b, x); }
				if (MODE == 12) { u[i] = prmt(u[i], 0x4B000000u, 0x7440 + (i & 3)); a[i].x = ffma(a[i].x, b.x, x.x); a[i].y = ffma(a[i].y, b.x, x.x); }
				if (MODE == 13) { u[i] = prmt(u[i], 0x4B000000u, 0x7440 + (i & 3)); a[i] = fadd2(a[i], b); }
				if (MODE == 14) { u[i] = u[i] * su + 7; }
				if (MODE == 16) { u[i] = dp4a_us(su ^ (unsigned)i, sv, u[i]); }
				if (MODE == 17) { u[i] = dp4a_us(su ^ (unsigned)i, sv, u[i]); a[i].x = ffma(a[i].x, b.x, x.x); }
				if (MODE == 18) { u[i] = dp4a_us(su ^ (unsigned)i, sv, u[i]); a[i] = ffma2(a[i], b, x); }
				if (MODE == 19) { u[i] = dp2a_us(su ^ (unsigned)i, sv, u[i]); }
				if (MODE == 20) { u[i] = dp4a_us(su ^ (unsigned)i, sv, u[i]); u[(i + 1) & 7] = prmt(u[(i + 1) & 7], 0x4B000000u, 0x7440 + (i & 3)); }
				if (MODE == 15) { u[i] = (u[i] ^ 0x80808080u) + (u[i] >> 3); }
			}
		}
	}
	float2 s = make_float2(0.f, 0.f);
#pragma unroll
	for (int i = 0; i < 8; i++) {
		s.x += a[i].x + __uint_as_float(u[i]);
		s.y += a[i].y;
	}
	out[blockIdx.x * 32 + threadIdx.x] = s;
}
template < int MODE > static void run(const char *name, float2 * d_out, double inst_per_iter)
{
	const int grid = 148 * 16, iters = 4000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k < MODE > <<<grid, 32 >>> (d_out, 10, make_float2(1.0001f, 0.9999f));
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	k < MODE > <<<grid, 32 >>> (d_out, iters, make_float2(1.0001f, 0.9999f));
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	int clk = 0;
	cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	const double cyc = ms * 1e-3 * clk * 1e3;
	const double winst_per_smsp = 4.0 * iters * inst_per_iter;	/* 4 warps per SMSP */
	printf("%-34s %8.3f ms  %.2f cycles per warp-instruction per SMSP (at %d MHz nominal)\n", name, ms, cyc / winst_per_smsp, clk / 1000);
}
int main()
{
	float2 *d_out;
	cudaMalloc(&d_out, sizeof(float2) * 148 * 16 * 32);
	run < 0 > ("FFMA2", d_out, 64);
	run < 1 > ("FADD2", d_out, 64);
	run < 2 > ("FFMA", d_out, 64);
	run < 3 > ("PRMT", d_out, 64);
	run < 4 > ("FFMA2+PRMT 1:1", d_out, 128);
	run < 5 > ("FFMA2+FADD2+PRMT 2:1:2", d_out, 160);
	run < 6 > ("FFMA+PRMT 1:1", d_out, 128);
	run < 7 > ("FFMA2 acc form (x*b+a)", d_out, 64);
	run < 8 > ("FFMA2 swapped A", d_out, 64);
	run < 9 > ("I2F.U8 byte-select", d_out, 64);
	run < 10 > ("I2F.U8 + 2 FFMA", d_out, 192);
	run < 11 > ("I2F.U8 + FFMA2", d_out, 128);
	run < 12 > ("PRMT + 2 FFMA", d_out, 192);
	run < 13 > ("PRMT + FADD2", d_out, 128);
	run < 14 > ("IMAD", d_out, 64);
	run < 16 > ("IDP4A", d_out, 64);
	run < 17 > ("IDP4A + FFMA", d_out, 128);
	run < 18 > ("IDP4A + FFMA2", d_out, 128);
	run < 19 > ("IDP2A", d_out, 64);
	run < 20 > ("IDP4A + PRMT", d_out, 128);
	run < 15 > ("LOP3+SHF+IADD (3 ALU)", d_out, 192);
	return 0;
}
