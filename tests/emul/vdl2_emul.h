/*
 * tests/emul/vdl2_emul.h -- TEST INFRASTRUCTURE: a 32-lane warp emulator for the host.
 *
 * vdlm2dec_b200/csrc/vdl2_demod.cuh (phase 2 of the kernel) is written against the tiny
 * `vw::` layer.  Under nvcc that layer is the real warp intrinsics; here each lane is a
 * ucontext fibre and every collective (shfl / ballot / sync) is a rendez-vous of all 32
 * fibres, so the SAME source runs on a CPU box and can be checked against the oracle
 * before any GPU time is spent.  Never part of the product library.
 */
#ifndef VDL2_EMUL_H
#define VDL2_EMUL_H
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector_types.h>
#include <vector_functions.h>

#define VQ static inline
#define VQ_COLD static
#define VQ_RARE static
#define VQ_PASSA static
#define VDL2_CONST

namespace vw {
extern int g_lane;		/* lane of the fibre that is running */
void barrier();			/* returns when all 32 lanes have arrived */
extern uint64_t g_xch[32];

VQ int lane() { return g_lane; }
VQ void sync() { barrier(); }
template <class T> VQ T shfl(T v, int src)
{
	static_assert(sizeof(T) <= 8, "shfl payload");
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	g_xch[g_lane] = raw;
	barrier();
	raw = g_xch[src & 31];
	barrier();
	T r;
	memcpy(&r, &raw, sizeof(T));
	return r;
}
template <class T> VQ T shfl_up(T v, int d)
{
	const int me = g_lane;
	return shfl(v, me - d >= 0 ? me - d : me);
}
VQ unsigned ballot(bool p)
{
	g_xch[g_lane] = p ? 1 : 0;
	barrier();
	unsigned m = 0;
	for (int i = 0; i < 32; i++)
		m |= (unsigned)(g_xch[i] & 1) << i;
	barrier();
	return m;
}
VQ unsigned atomic_inc(unsigned *p) { return (*p)++; }
VQ void fence() {}
VQ int ffs(unsigned m) { return __builtin_ffs((int)m); }
VQ int popc(unsigned m) { return __builtin_popcount(m); }
VQ void sincos(float a, float &s, float &c) { sincosf(a, &s, &c); }
VQ float rsqrt(float a) { return 1.0f / sqrtf(a); }
VQ float fma(float a, float b, float c) { return fmaf(a, b, c); }
VQ float atan2(float y, float x) { return atan2f(y, x); }
VQ float fdiv(float a, float b) { return a / b; }
VQ float fsub(float a, float b) { return a - b; }
VQ float fadd(float a, float b) { return a + b; }
VQ float fmul(float a, float b) { return a * b; }
template <class T> VQ T ldcg(const T * p) { return *p; }
template <class T> VQ T ldg(const T * p) { return *p; }
template <class T> VQ T *as_shared(T * p) { return p; }
VQ unsigned f2bits(float a) { unsigned u; memcpy(&u, &a, 4); return u; }
VQ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
VQ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
}
#endif
