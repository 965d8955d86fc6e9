/*
 * oracle/ref/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference demodulator *in place*: this translation unit
 * `#include`s /root/reference/d8psk.c (found through -I at build time, see
 * oracle/Makefile), so its `static inline` functions (demodD8psk, filteredphase,
 * putbit ...) and `channel_t` are visible here without copying a line of it.
 * Outputs go to oracle/_ref/ (git-ignored; they travel to the GPU box as binaries).
 *
 * What this file adds around the reference code:
 *   - a zero-initialised channel_t (bug 1: the reference leaves
 *     clk/Inbuff/Ph/p2err uninitialised on the rcv_thread stack, d8psk.c:338),
 *   - the 20-line sample loop of rcv_thread (d8psk.c:343-382) driven from a caller
 *     buffer instead of the global Cbuff + barriers (state carried across calls the
 *     same way the never-returning thread carries its locals),
 *   - capturing replacements for initVdlm2()/decodeVdlm2() (vdlm2.c:163-206) so the
 *     completed msgblk_t is observed at the hand-off boundary (tap T6),
 *   - taps T1..T5 derived from channel_t before/after each demodD8psk() call.
 *
 * The sample conversion of rtl.c:285-292 (u8 - 127.37f, including its index quirk)
 * is restated in orc_feed_cu8() because rtl.c cannot be compiled without librtlsdr.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <complex.h>
#include <pthread.h>
#include <time.h>

/* ---- globals the reference translation unit imports (vdlm2.h:30-31,85-99) ---- */
unsigned int SDRINRATE = 2000000;
unsigned int SDRCLK = 500;
unsigned int Fc = 0;
int ppm = 0;
int verbose = 0;
FILE *logfd = NULL;
pthread_barrier_t Bar1, Bar2;
int grndmess, emptymess, undecmess;
#ifndef WITH_RTL
#define WITH_RTL 1
#endif

/* the reference demodulator, verbatim, from the read-only mount (it includes vdlm2.h,
   which has no include guard, so nothing else here may include it) */
#include "d8psk.c"
complex float Cbuff[RTLINBUFSZ / 2];

#include "../orc_api.h"

typedef struct {
	channel_t ch;
	/* mixer state: the locals of rcv_thread, d8psk.c:343-347 */
	int clk, nf, no, nwf;
	complex float D;
	complex float *wf;
	int real_input;
	unsigned fs, sdrclk;
	uint32_t taps;
	int64_t ndump;		/* dumps produced so far */
	int64_t nsamp;
	int64_t sync_dump;	/* dump index of the last trigger */
	/* taps */
	orc_vec dumps, steps, syncs, syms, blocks;
} refctx;

static refctx *g_cur;		/* context running inside demodD8psk (single threaded per feed) */
static pthread_mutex_t g_mtx = PTHREAD_MUTEX_INITIALIZER;	/* viterbi.c has global state */

static void vec_push(orc_vec * v, const void *rec, size_t sz)
{
	if ((v->n + 1) * sz > v->cap) {
		v->cap = v->cap ? v->cap * 2 : 64 * sz;
		if (v->cap < (v->n + 1) * sz)
			v->cap = (v->n + 1) * sz;
		v->p = realloc(v->p, v->cap);
	}
	memcpy((char *)v->p + v->n * sz, rec, sz);
	v->n++;
}

/* ---- capturing stand-ins for vdlm2.c:163-206 (the block queue is downstream of the path) ---- */
int initVdlm2(channel_t * ch)
{
	ch->state = WSYNC;
	ch->blk = calloc(sizeof(msgblk_t), 1);
	ch->blk->chn = ch->chn;
	ch->blk->Fr = ch->Fr;
	return 0;
}

void stopVdlm2(void)
{
}

void decodeVdlm2(channel_t * ch)
{
	refctx *c = g_cur;
	if (c && (c->taps & ORC_TAP_BLOCKS)) {
		orc_block b;
		memset(&b, 0, sizeof(b));
		b.sync_dump = c->sync_dump;
		b.end_dump = c->ndump - 1;
		b.chn = ch->blk->chn;
		b.Fr = ch->blk->Fr;
		b.ppm = ch->blk->ppm;
		b.nbrow = ch->blk->nbrow;
		b.nlbyte = ch->blk->nlbyte;
		for (int r = 0; r < 8; r++)
			memcpy(b.data[r], ch->blk->data[r], 255);
		vec_push(&c->blocks, &b, sizeof(b));
	}
	free(ch->blk);
	ch->blk = calloc(sizeof(msgblk_t), 1);
	ch->blk->chn = ch->chn;
	ch->blk->Fr = ch->Fr;
}

void *orc_open(int chn, int Fr, int Fo, unsigned fs, unsigned sdrclk, int real_input, uint32_t taps)
{
	refctx *c = calloc(1, sizeof(refctx));	/* zero-init: see header comment */
	c->ch.chn = chn;
	c->ch.Fr = Fr;
	c->fs = fs;
	c->sdrclk = sdrclk;
	c->taps = taps;
	c->real_input = real_input;
	c->sync_dump = -1;
	initD8psk(&c->ch);
	initVdlm2(&c->ch);
	/* local oscillator exactly as d8psk.c:353-357 (SDRINRATE is read through the global) */
	SDRINRATE = fs;
	SDRCLK = sdrclk;
	c->nwf = fs / STEPRATE;
	c->wf = malloc(sizeof(complex float) * c->nwf);
	{
		float Fof;
		int no;
		Fof = (float)Fo / (float)(SDRINRATE) * 2.0 * M_PI;
		for (no = 0; no < (int)(SDRINRATE / STEPRATE); no++)
			c->wf[no] = cexpf(-no * Fof * I);
	}
	return c;
}

void orc_close(void *h)
{
	refctx *c = h;
	if (!c)
		return;
	free(c->ch.blk);
	free(c->wf);
	free(c->dumps.p);
	free(c->steps.p);
	free(c->syncs.p);
	free(c->syms.p);
	free(c->blocks.p);
	free(c);
}

/* one decimated sample into the reference demodulator, with taps around the call */
static inline void one_dump(refctx * c, complex float D)
{
	channel_t *ch = &c->ch;
	if (c->taps == 0) {
		demodD8psk(ch, D);
		c->ndump++;
		return;
	}
	int st0 = ch->state;
	int clk0 = ch->clk;
	float P1_0 = ch->P1, df0 = ch->df;
	int phidx0 = ch->Phidx;

	if (c->taps & ORC_TAP_DUMPS) {
		float d[2] = { crealf(D), cimagf(D) };
		vec_push(&c->dumps, d, sizeof(d));
	}
	int64_t idx = c->ndump++;
	demodD8psk(ch, D);
	int trig = (st0 == WSYNC && ch->state != WSYNC);
	if (trig)
		c->sync_dump = idx;	/* blocks are stamped with the dump index of their trigger */

	if (st0 == WSYNC) {
		if (ch->Phidx != phidx0) {	/* a WSYNC step ran (d8psk.c:252-313) */
			if (c->taps & ORC_TAP_STEPS) {
				orc_step s;
				s.pad = 0;
				s.dump = c->ndump - 1;
				s.P = ch->Ph[ch->Phidx];
				s.err = trig ? -1.0f : ch->perr;
				s.fr = trig ? ch->df : ch->pfr;
				vec_push(&c->steps, &s, sizeof(s));
			}
			if (trig && (c->taps & ORC_TAP_SYNCS)) {
				orc_sync s;
				s.dump = c->ndump - 1;
				s.clk = ch->clk;
				s.df = ch->df;
				s.ppm = ch->blk->ppm;
				s.P1 = ch->P1;
				vec_push(&c->syncs, &s, sizeof(s));
			}
		}
	} else if (clk0 + 4 >= 32) {	/* a data symbol was sliced (d8psk.c:317-331) */
		if (c->taps & ORC_TAP_SYMS) {
			/* D is a local of demodD8psk; recompute it with the same operations from
			   P1 before/after and df (d8psk.c:323-327) */
			float P = ch->P1;
			float Dd = (P - P1_0) - df0;
			if (Dd > M_PI)
				Dd -= 2 * M_PI;
			if (Dd < -M_PI)
				Dd += 2 * M_PI;
			int gi = (int)roundf(128.0 * Dd / M_PI + 128.0);	/* d8psk.c:213 */
			orc_sym s;
			memset(&s, 0, sizeof s);
			s.dump = c->ndump - 1;
			s.D = Dd;
			s.P = P;
			s.gi = gi;
			s.v[0] = Grey1[gi];
			s.v[1] = Grey2[gi];
			s.v[2] = Grey3[gi];
			s.state_after = ch->state;
			vec_push(&c->syms, &s, sizeof(s));
		}
	}
}

#define MIX_LOOP(SAMPLE_EXPR, N)                                           \
	for (size_t i = 0; i < (N); i++) {                                   \
		c->D += (SAMPLE_EXPR) * c->wf[c->no];                      \
		c->nf++;                                                   \
		c->no = (c->no + 1) % c->nwf;                              \
		c->clk += 21;                                              \
		if (c->clk >= (int)c->sdrclk) {                            \
			c->clk %= c->sdrclk;                               \
			c->D /= c->nf;                                     \
			one_dump(c, c->D);                                 \
			c->D = 0;                                          \
			c->nf = 0;                                         \
		}                                                          \
	}                                                                  \
	c->nsamp += (N);

void orc_feed_cf32(void *h, const float *iq, size_t n)
{
	refctx *c = h;
	const complex float *x = (const complex float *)iq;
	pthread_mutex_lock(&g_mtx);
	g_cur = c;
	SDRINRATE = c->fs;
	SDRCLK = c->sdrclk;
	MIX_LOOP(x[i], n);
	g_cur = NULL;
	pthread_mutex_unlock(&g_mtx);
}

void orc_feed_f32real(void *h, const float *xr, size_t n)
{
	refctx *c = h;
	pthread_mutex_lock(&g_mtx);
	g_cur = c;
	SDRINRATE = c->fs;
	SDRCLK = c->sdrclk;
	MIX_LOOP(xr[i], n);
	g_cur = NULL;
	pthread_mutex_unlock(&g_mtx);
}

/* rtl.c:285-292 restated: r = (float)u8 - (float)127.37 ; `offset` lets cs8-style
   variants (offset 0 on a signed byte) share the path. */
void orc_feed_cu8(void *h, const uint8_t * iq, size_t n, float offset)
{
	enum { CH = 4096 };
	float buf[2 * CH];
	while (n) {
		size_t m = n < CH ? n : CH;
		for (size_t k = 0; k < m; k++) {
			buf[2 * k] = (float)iq[2 * k] - offset;
			buf[2 * k + 1] = (float)iq[2 * k + 1] - offset;
		}
		orc_feed_cf32(h, buf, m);
		iq += 2 * m;
		n -= m;
	}
}

void orc_feed_cs8(void *h, const int8_t * iq, size_t n)
{
	enum { CH = 4096 };
	float buf[2 * CH];
	while (n) {
		size_t m = n < CH ? n : CH;
		for (size_t k = 0; k < 2 * m; k++)
			buf[k] = (float)iq[k];
		orc_feed_cf32(h, buf, m);
		iq += 2 * m;
		n -= m;
	}
}

void orc_feed_cs16(void *h, const int16_t * iq, size_t n)
{
	enum { CH = 4096 };
	float buf[2 * CH];
	while (n) {
		size_t m = n < CH ? n : CH;
		for (size_t k = 0; k < 2 * m; k++)
			buf[k] = (float)iq[k];
		orc_feed_cf32(h, buf, m);
		iq += 2 * m;
		n -= m;
	}
}

/* the block-level quirk of rtl.c:285-292: `Cbuff[i / 2]` is indexed AFTER both
   increments, so sample k of a 32768-sample block lands in slot k+1, slot 0 keeps
   its previous content and the last sample is written out of bounds (dropped here).
   `blk` must hold exactly RTLINBUFSZ bytes. */
void orc_feed_rtl_block_quirk(void *h, const uint8_t * blk)
{
	static __thread float buf[RTLINBUFSZ + 2];
	refctx *c = h;
	(void)c;
	/* slot 0 is never written by the callback: it is the zero of the static array */
	buf[0] = 0;
	buf[1] = 0;
	for (int k = 0; k < RTLINBUFSZ / 2 - 1; k++) {
		buf[2 * (k + 1)] = (float)blk[2 * k] - (float)127.37;
		buf[2 * (k + 1) + 1] = (float)blk[2 * k + 1] - (float)127.37;
	}
	orc_feed_cf32(h, buf, RTLINBUFSZ / 2);
}

const void *orc_tap(void *h, int which, size_t *count)
{
	refctx *c = h;
	orc_vec *v = NULL;
	switch (which) {
	case ORC_TAP_DUMPS: v = &c->dumps; break;
	case ORC_TAP_STEPS: v = &c->steps; break;
	case ORC_TAP_SYNCS: v = &c->syncs; break;
	case ORC_TAP_SYMS: v = &c->syms; break;
	case ORC_TAP_BLOCKS: v = &c->blocks; break;
	}
	if (!v) {
		*count = 0;
		return NULL;
	}
	*count = v->n;
	return v->p;
}

void orc_clear_taps(void *h)
{
	refctx *c = h;
	c->dumps.n = c->steps.n = c->syncs.n = c->syms.n = c->blocks.n = 0;
}

int64_t orc_ndump(void *h)
{
	return ((refctx *) h)->ndump;
}

const char *orc_kind(void)
{
	return "reference";
}

/* the reference's own constant tables (d8psk.h), for the known-answer tests */
/* the oscillator table of an open handle, exactly as d8psk.c:353-357 builds it (cexpf): n complex floats -> out[2n] */
int orc_nco(void *h, float *out, int max)
{
	refctx *c = h;
	int n = c->nwf < max ? c->nwf : max;
	for (int i = 0; i < n; i++) {
		out[2 * i] = crealf(c->wf[i]);
		out[2 * i + 1] = cimagf(c->wf[i]);
	}
	return c->nwf;
}

const float *orc_table(int which)
{
	switch (which) {
	case 0: return SW;
	case 1: return mflt;
	case 2: return Grey1;
	case 3: return Grey2;
	case 4: return Grey3;
	}
	return NULL;
}

/* Timing leg (bench.py cpu_baseline / --impl reference): one private channel per
   calling thread, taps off, `reps` passes over the caller's cu8 buffer through the
   restated rtl.c conversion + rcv_thread loop.  Returns seconds (CLOCK_MONOTONIC).
   The viterbi.c globals are shared between threads exactly as in the reference
   (vdlm2.h:124-126 / viterbi.c:25-27), so no lock is taken here. */
double orc_time_cu8(int Fr, int Fo, unsigned fs, unsigned sdrclk, const uint8_t * iq, size_t n, int reps)
{
	refctx *c = orc_open(0, Fr, Fo, fs, sdrclk, 0, 0);
	enum { CH = 32768 };
	complex float *buf = malloc(sizeof(complex float) * CH);
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int r = 0; r < reps; r++) {
		size_t off = 0;
		while (off < n) {
			size_t m = (n - off) < CH ? (n - off) : CH;
			for (size_t k = 0; k < m; k++) {
				float re = (float)iq[2 * (off + k)] - (float)127.37;
				float im = (float)iq[2 * (off + k) + 1] - (float)127.37;
				buf[k] = re + im * I;
			}
			MIX_LOOP(buf[i], m);
			off += m;
		}
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(buf);
	orc_close(c);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
