/*
 * d8psk_shim.c -- drop-in replacement for the reference object d8psk.o.
 *
 * Exports exactly the symbols d8psk.c exports (vdlm2.h:113-114,128):
 *     int initD8psk(channel_t *), void *rcv_thread(void *), unsigned reversebits(unsigned, int)
 * and imports what d8psk.c imports from the untouched host files: Cbuff, SDRINRATE, SDRCLK,
 * Bar1, Bar2, nbch (main.c:59), initVdlm2(), decodeVdlm2() (vdlm2.c:163-206).
 * It is compiled against the reference's own vdlm2.h (-I<reference>) -- nothing of the
 * reference is copied here -- and linked with libvdl2gpu.so.  See INTEGRATION.md.
 *
 * Protocol (d8psk.c:335-385): main.c:228-231 starts one rcv_thread per channel, never joins
 * it.  Every thread registers its (chn, Fr, Fo) and its channel_t, calls initVdlm2() exactly
 * like the reference (channel 0 thereby starts blk_thread), and keeps the two-barrier
 * lock-step with the SDR callback (rtl.c:283,294).  The thread of channel 0 owns the GPU
 * handle: after Bar2 it hands the whole converted block Cbuff (complex float[32768],
 * vdlm2.h:89) to the fused kernel -- all channels are demodulated from that one stream --
 * drains the completed blocks and passes each to decodeVdlm2() through the channel_t of
 * the channel it belongs to (ownership of ch->blk as in vdlm2.c:189-205).
 * CUDA failure: message on stderr + exit(1) (the reference's only error convention).
 *
 * -DVDL2_SHIM_LINK (the next row of the scope table): the object ALSO takes the
 * place of vdlm2.o and rs.o.  It then exports initVdlm2 / decodeVdlm2 / stopVdlm2 (vdlm2.h:116-118) itself,
 * no blk_thread is started, and instead of handing blocks to decodeVdlm2() the leader drains FRAMES -- the
 * blocks went through rs(), HDLC un-stuffing and the FCS check on the device (vdl2_drain_frames) -- and calls
 * out(blk, hdata, l) (vdlm2.h:134) exactly where check_frame() would (vdlm2.c:60).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "vdlm2.h"		/* the reference's header, found through -I */
#include "vdl2gpu.h"

extern int nbch;		/* main.c:59 */

static channel_t *g_ch[MAXNBCHANNELS];
static vdl2_chan_param_t g_par[MAXNBCHANNELS];
static vdl2gpu_t *g_gpu;
static pthread_mutex_t g_busy = PTHREAD_MUTEX_INITIALIZER;

/* main.c:246 calls exit(1) as soon as the SDR stops, while workers may still be inside the last
   block.  Registered after the CUDA runtime's own atexit hook (so it runs before it): wait for the
   block in flight instead of tearing the context down under it. */
static void quiesce(void)
{
	pthread_mutex_lock(&g_busy);
}

int initD8psk(channel_t * ch)
{				/* same fields the reference sets (d8psk.c:28-37); the demodulator state itself lives on the GPU */
	ch->ink = 0;
	ch->Phidx = 0;
	ch->df = 0;
	ch->perr = 100;
	ch->P1 = 0;
	return 0;
}

unsigned int reversebits(const unsigned int bits, const int n)
{				/* also used by out.c:429-432 */
	unsigned int in = bits, out = 0;
	for (int i = 0; i < n; i++) {
		out = (out << 1) | (in & 1);
		in >>= 1;
	}
	return out;
}

static void die(const char *what)
{
	fprintf(stderr, "vdl2gpu shim: %s: %s\n", what, vdl2_last_error(g_gpu));
	exit(1);
}

static void gpu_open(void)
{
	vdl2_config_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.fs = SDRINRATE;
	cfg.sdrclk = SDRCLK;
#ifdef WITH_AIR
	cfg.format = VDL2_FMT_F32REAL;
#else
	cfg.format = VDL2_FMT_CF32;
#endif
	cfg.nch = nbch;
	cfg.ch_per_stream = nbch;
	cfg.device = getenv("VDL2_GPU_DEVICE") ? atoi(getenv("VDL2_GPU_DEVICE")) : 0;
	cfg.max_samples = RTLINBUFSZ / 2 + SDRINRATE / 1000;
	cfg.max_blocks = 1024;
	if (vdl2_create(&cfg, g_par, &g_gpu))
		die("vdl2_create");
	atexit(quiesce);
}

#ifdef VDL2_SHIM_LINK
int initVdlm2(channel_t * ch)
{				/* vdlm2.c:163-180 without the consumer thread: frames are produced on the device */
	ch->state = WSYNC;
	ch->blk = calloc(sizeof(msgblk_t), 1);
	ch->blk->chn = ch->chn;
	ch->blk->Fr = ch->Fr;
	return 0;
}

void decodeVdlm2(channel_t * ch)
{				/* never called by this object; kept for link compatibility (vdlm2.h:118) */
	(void)ch;
}

void stopVdlm2(void)
{				/* main.c:108,244: nothing is queued on the host */
}

static void gpu_block(void)
{
	static vdl2_frame_t fr[1024];
	int nf = 0, nb = 0;
	pthread_mutex_lock(&g_busy);
	if (vdl2_process_host(g_gpu, Cbuff, RTLINBUFSZ / 2, 0))
		die("vdl2_process_host");
	if (vdl2_drain_frames(g_gpu, fr, 1024, &nf, NULL, 0, &nb))	/* out*.c reads chn, Fr, ppm, tv only (out.c:169-230,543) */
		die("vdl2_drain_frames");
	for (int i = 0; i < nf; i++) {
		msgblk_t blk;	/* what check_frame() passes on (vdlm2.c:60): only the header fields are read downstream */
		memset(&blk, 0, sizeof blk);
		blk.chn = fr[i].chn;
		blk.Fr = fr[i].Fr;
		blk.ppm = fr[i].ppm;
		gettimeofday(&blk.tv, NULL);	/* d8psk.c:295 (wall clock in the reference too) */
		out(&blk, fr[i].hdata, fr[i].len);
	}
	pthread_mutex_unlock(&g_busy);
}
#else
static void gpu_block(void)
{
	static vdl2_block_t out[1024];
	int n = 0;
	pthread_mutex_lock(&g_busy);
	if (vdl2_process_host(g_gpu, Cbuff, RTLINBUFSZ / 2, 0))
		die("vdl2_process_host");
	if (vdl2_drain_blocks(g_gpu, out, 1024, &n))
		die("vdl2_drain_blocks");
	for (int i = 0; i < n; i++) {
		channel_t *ch = NULL;
		for (int c = 0; c < nbch; c++)
			if (g_ch[c] && g_ch[c]->chn == out[i].chn)
				ch = g_ch[c];
		if (!ch)
			continue;
		msgblk_t *blk = ch->blk;
		gettimeofday(&blk->tv, NULL);	/* d8psk.c:295 (wall clock in the reference too) */
		blk->ppm = out[i].ppm;
		blk->nbrow = out[i].nbrow;
		blk->nlbyte = out[i].nlbyte;
		for (int r = 0; r < 8; r++)
			memcpy(blk->data[r], out[i].data[r], 255);
		decodeVdlm2(ch);	/* takes blk, installs a fresh zeroed one (vdlm2.c:189-205) */
	}
	pthread_mutex_unlock(&g_busy);
}
#endif

void *rcv_thread(void *arg)
{
	thread_param_t *param = (thread_param_t *) arg;
	channel_t *ch = calloc(1, sizeof(channel_t));
	ch->chn = param->chn;
	ch->Fr = param->Fr;
	initD8psk(ch);
	initVdlm2(ch);
	g_par[param->chn].chn = param->chn;
	g_par[param->chn].Fr = param->Fr;
	g_par[param->chn].Fo = param->Fo;
	g_ch[param->chn] = ch;

	pthread_barrier_wait(&Bar1);	/* all nbch workers have registered once this returns */
	const int leader = (param->chn == 0);
	if (leader)
		gpu_open();
	do {
		pthread_barrier_wait(&Bar2);
		if (leader)
			gpu_block();
		pthread_barrier_wait(&Bar1);
	} while (1);
	return NULL;
}
