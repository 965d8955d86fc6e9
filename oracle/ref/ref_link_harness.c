/*
 * oracle/ref/ref_link_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED link layer of the reference *in place*: this translation unit `#include`s
 * /root/reference/vdlm2.c (found through -I, see oracle/Makefile) and is linked with the reference's
 * rs.c and crc.c, so blk_thread / check_frame / set_eras (vdlm2.c:40-161) and rs() (rs.c:81-291) run
 * exactly as shipped.  Nothing of the reference is copied.
 *
 * What this file adds around it:
 *   - out() (out.c:426, the consumer of check_frame) replaced by a capture of (hdata, l);
 *   - rs() reached through a wrapper that records its return value (the reference ignores it);
 *   - free() of a finished block (vdlm2.c:157) hooked: the moment blk_thread is done with a block is the
 *     only completion signal the reference has; the rows after rs() are captured there;
 *   - blocks are pushed through initVdlm2()/decodeVdlm2() (vdlm2.c:163-206) one at a time.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <stdio.h>
#include <unistd.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#include <pthread.h>
#include <time.h>

int verbose = 0;
FILE *logfd = NULL;

static void harness_free(void *p);
static int harness_rs(unsigned char *data, int *eras_pos, int no_eras);
#define free(p) harness_free(p)
#define rs(a, b, c) harness_rs(a, b, c)
#ifndef WITH_RTL
#define WITH_RTL 1
#endif
#include "vdlm2.c"		/* the reference link layer, verbatim, from the read-only mount */
#undef free
#undef rs
extern int rs(unsigned char *data, int *eras_pos, int no_eras);	/* rs.c:81 */

#include "../orc_link_api.h"

static pthread_mutex_t h_mtx = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t h_cnd = PTHREAD_COND_INITIALIZER;
static long h_done;
static channel_t h_ch;
static int h_started;
/* capture state of the block in flight (one at a time) */
static orc_frame *c_frames;
static int c_max, c_n, c_block, c_overflow, c_row;
static orc_blkstat *c_stat;
static uint8_t *c_rows;
static const orc_block *c_src;

void out(msgblk_t * blk, unsigned char *hdata, int l)
{				/* stands in for out.c:426 */
	(void)blk;
	if (c_stat)
		c_stat->nframes++;
	if (!c_frames)
		return;
	if (c_n >= c_max) {
		c_overflow = 1;
		return;
	}
	orc_frame *f = &c_frames[c_n++];
	memset(f, 0, sizeof *f);
	f->block = c_block;
	f->len = l;
	f->chn = c_src->chn;
	f->Fr = c_src->Fr;
	f->ppm = c_src->ppm;
	f->sync_dump = c_src->sync_dump;
	memcpy(f->hdata, hdata, l < ORC_FRAME_MAX ? l : ORC_FRAME_MAX);
}

static int harness_rs(unsigned char *data, int *eras_pos, int no_eras)
{
	const int r = rs(data, eras_pos, no_eras);
	if (c_stat && c_row < 8)
		c_stat->rs[c_row] = (int8_t) r;
	c_row++;
	return r;
}

static void harness_free(void *p)
{				/* vdlm2.c:157: blk_thread is done with this block */
	msgblk_t *blk = (msgblk_t *) p;
	if (c_rows)
		for (int r = 0; r < 8; r++)
			memcpy(c_rows + r * 255, blk->data[r], 255);
	free(p);
	pthread_mutex_lock(&h_mtx);
	h_done++;
	pthread_cond_broadcast(&h_cnd);
	pthread_mutex_unlock(&h_mtx);
}

static void push_block(const orc_block * b)
{
	if (!h_started) {
		memset(&h_ch, 0, sizeof h_ch);
		h_ch.chn = 0;	/* channel 0 starts blk_thread (vdlm2.c:171-176) */
		initVdlm2(&h_ch);
		h_started = 1;
	}
	msgblk_t *blk = h_ch.blk;	/* zeroed by calloc (vdlm2.c:168,201) */
	blk->chn = b->chn;
	blk->Fr = b->Fr;
	blk->ppm = b->ppm;
	blk->nbrow = b->nbrow;
	blk->nlbyte = b->nlbyte;
	for (int r = 0; r < 8; r++)
		memcpy(blk->data[r], b->data[r], 255);
	pthread_mutex_lock(&h_mtx);
	const long want = h_done + 1;
	pthread_mutex_unlock(&h_mtx);
	decodeVdlm2(&h_ch);	/* hands the block to blk_thread, installs a fresh one (vdlm2.c:189-205) */
	pthread_mutex_lock(&h_mtx);
	while (h_done < want)
		pthread_cond_wait(&h_cnd, &h_mtx);
	pthread_mutex_unlock(&h_mtx);
}

int orc_link_decode(const orc_block * blocks, int n, orc_frame * frames, int max_frames, int *n_frames, orc_blkstat * stats,
		    uint8_t * rows_after)
{
	static pthread_mutex_t api = PTHREAD_MUTEX_INITIALIZER;
	pthread_mutex_lock(&api);
	c_frames = frames;
	c_max = max_frames;
	c_n = 0;
	c_overflow = 0;
	for (int i = 0; i < n; i++) {
		c_block = i;
		c_src = &blocks[i];
		c_row = 0;
		c_stat = stats ? &stats[i] : NULL;
		if (c_stat)
			memset(c_stat, 0, sizeof *c_stat);
		c_rows = rows_after ? rows_after + (size_t) i * 8 * 255 : NULL;
		push_block(&blocks[i]);
	}
	if (n_frames)
		*n_frames = c_n;
	c_frames = NULL;
	c_stat = NULL;
	c_rows = NULL;
	const int ov = c_overflow;
	pthread_mutex_unlock(&api);
	return ov;
}

double orc_link_time(const orc_block * blocks, int n, int reps)
{
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int r = 0; r < reps; r++)
		orc_link_decode(blocks, n, NULL, 0, NULL, NULL, NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

const char *orc_link_kind(void)
{
	return "reference vdlm2.c + rs.c + crc.c compiled in place";
}
