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
 *
 * -DVDL2_SHIM_FILE (row f2, file replay): the object is linked next to file_shim.o, which takes the place of
 * rtl.o.  There is no SDR callback and no Cbuff then: the workers only register their channel and park on the
 * barrier; file_shim.c feeds raw captures through vdl2shim_open() / vdl2shim_feed() below (shim_internal.h).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>
#include "vdlm2.h"		/* the reference's header, found through -I */
#include "vdl2gpu.h"
#include "shim_internal.h"

extern int nbch;		/* main.c:59 */

static channel_t *g_ch[MAXNBCHANNELS];
static vdl2_chan_param_t g_par[MAXNBCHANNELS];
static vdl2gpu_t *g_gpu;
static pthread_mutex_t g_busy = PTHREAD_ERRORCHECK_MUTEX_INITIALIZER_NP;

/* main.c:246 calls exit(1) as soon as the SDR stops, while workers may still be inside the last
   block.  Registered after the CUDA runtime's own atexit hook (so it runs before it): wait for the
   block in flight instead of tearing the context down under it.  The mutex is error-checking because
   main.c's signal handler (main.c:106-110) calls exit() on the main thread, which in the replay build is
   the very thread that feeds the GPU: it must not wait for itself (EDEADLK instead of a hang on Ctrl-C). */
static void deliver(void);
#ifdef VDL2_SHIM_LINK
static void out_drain_wait(void);
#endif
static int g_last_n;		/* blocks handed over by the last deliver() */
static double g_first = -1.0, g_last;	/* VDL2_SHIM_STATS: wall clock of the first feed and of the end of the last one */
static unsigned long long g_fed;

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

static void quiesce(void)
{
	if (pthread_mutex_lock(&g_busy) != 0)
		return;		/* exit() from the signal handler on the thread that is feeding: nothing can be collected safely */
	if (!g_gpu)
		return;
	/* feeds are asynchronous: what the last rounds completed is still on the device.  Collect it, and give the reference's
	   consumer thread the time to print it before the process goes away (its stopVdlm2() has already returned). */
	deliver();
#ifdef VDL2_SHIM_LINK
	out_drain_wait();	/* the consumer thread has printed everything */
#endif
	g_last = now();
	if (getenv("VDL2_SHIM_STATS") && g_first >= 0)
		fprintf(stderr, "vdl2gpu shim: fed %llu samples in %.6f s (first feed to last block delivered)\n", g_fed, g_last - g_first);
	if (g_last_n)
		usleep(100000 + 500 * (unsigned)g_last_n);
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

int vdl2shim_nch(void)
{
	return nbch;
}

/* blocks / frames pending between two drains never exceed the queue, so one drain call always has room */
#define SHIM_QCAP 4096
static vdl2_block_t *g_blocks;
#define SHIM_BYTES ((size_t) SHIM_QCAP * 512)
static vdl2_frame_hdr_t *g_hdrs;	/* VDL2_SHIM_LINK: packed frames (page-locked) */
static uint8_t *g_bytes, g_one[2048 + 64];
static double g_t0 = -1;	/* VDL2_FILE_T0: epoch of sample 0 of a replayed capture; < 0 = wall clock as in d8psk.c:295 */

void vdl2shim_open(unsigned fs, unsigned sdrclk, int format, size_t max_samples)
{
	vdl2_config_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.fs = fs;
	cfg.sdrclk = sdrclk;
	cfg.format = format;
	cfg.nch = nbch;
	cfg.ch_per_stream = nbch;
	cfg.device = getenv("VDL2_GPU_DEVICE") ? atoi(getenv("VDL2_GPU_DEVICE")) : 0;
	cfg.max_samples = max_samples + fs / 1000;	/* room for the carried sub-millisecond tail */
	cfg.max_blocks = SHIM_QCAP;
	if (vdl2_create(&cfg, g_par, &g_gpu))
		die("vdl2_create");
	g_blocks = malloc(sizeof(vdl2_block_t) * SHIM_QCAP);
	if (!g_blocks || vdl2_host_alloc(sizeof(vdl2_frame_hdr_t) * SHIM_QCAP, (void **)&g_hdrs) || vdl2_host_alloc(SHIM_BYTES, (void **)&g_bytes)) {
		fprintf(stderr, "vdl2gpu shim: out of memory\n");
		exit(1);
	}
	if (getenv("VDL2_FILE_T0"))
		g_t0 = atof(getenv("VDL2_FILE_T0"));
	atexit(quiesce);
}

static void stamp(struct timeval *tv, int64_t sync_dump)
{				/* d8psk.c:295 takes the wall clock at the trigger; a replay can ask for capture time instead */
	if (g_t0 < 0) {
		gettimeofday(tv, NULL);
		return;
	}
	const double t = g_t0 + (double)sync_dump / 84000.0;	/* 8 x 10500 decimated samples per second (d8psk.h) */
	tv->tv_sec = (time_t) t;
	tv->tv_usec = (suseconds_t) ((t - (double)tv->tv_sec) * 1e6);
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

/* out() runs on its own consumer thread, like the reference's blk_thread (vdlm2.c:84-161, the sole caller of out()): the feeder
   only copies the packed frames of a drain into a queue entry, so formatting overlaps the next launch instead of delaying it */
struct outq_entry {
	struct outq_entry *next;
	int nf;
	vdl2_frame_hdr_t *hdrs;
	uint8_t *bytes;
};
static struct outq_entry *q_head, *q_tail;
static pthread_mutex_t q_mtx = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t q_cond = PTHREAD_COND_INITIALIZER, q_idle = PTHREAD_COND_INITIALIZER;
static int q_busy, q_started;

static void *out_thread(void *arg)
{
	(void)arg;
	for (;;) {
		pthread_mutex_lock(&q_mtx);
		while (!q_head)
			pthread_cond_wait(&q_cond, &q_mtx);
		struct outq_entry *e = q_head;
		q_head = e->next;
		if (!q_head)
			q_tail = NULL;
		q_busy = 1;
		pthread_mutex_unlock(&q_mtx);
		for (int i = 0; i < e->nf; i++) {
			msgblk_t blk;	/* what check_frame() passes on (vdlm2.c:60): only the header fields are read downstream */
			memset(&blk, 0, sizeof blk);
			blk.chn = e->hdrs[i].chn;
			blk.Fr = e->hdrs[i].Fr;
			blk.ppm = e->hdrs[i].ppm;
			stamp(&blk.tv, e->hdrs[i].sync_dump);
			/* out() -> outacars() strips parity bits IN PLACE (outacars.c:223-226) and reads a little past l: own scratch copy */
			memset(g_one, 0, sizeof g_one);
			memcpy(g_one, e->bytes + e->hdrs[i].offset, (size_t) e->hdrs[i].len);
			out(&blk, g_one, e->hdrs[i].len);
		}
		free(e->hdrs);
		free(e->bytes);
		free(e);
		pthread_mutex_lock(&q_mtx);
		q_busy = 0;
		if (!q_head)
			pthread_cond_broadcast(&q_idle);
		pthread_mutex_unlock(&q_mtx);
	}
	return NULL;
}

static void out_drain_wait(void)
{
	pthread_mutex_lock(&q_mtx);
	while (q_head || q_busy)
		pthread_cond_wait(&q_idle, &q_mtx);
	pthread_mutex_unlock(&q_mtx);
}

static void deliver(void)
{				/* called with g_busy held, after a process call returned */
	/* frames leave the device ordered (completion order, the order blk_thread would see the blocks) and packed: a 32-byte
	   header + the frame's own bytes instead of 2048-byte records (vdl2_drain_frames_packed; page-locked buffers) */
	int nf = 0;
	size_t nb = 0;
	if (vdl2_drain_frames_packed(g_gpu, g_hdrs, SHIM_QCAP, &nf, g_bytes, SHIM_BYTES, &nb, NULL))	/* out*.c reads chn, Fr, ppm, tv only (out.c:169-230,543) */
		die("vdl2_drain_frames_packed");
	if (nf == 0)
		return;
	struct outq_entry *e = malloc(sizeof *e);
	e->next = NULL;
	e->nf = nf;
	e->hdrs = malloc(sizeof(vdl2_frame_hdr_t) * (size_t) nf);
	e->bytes = malloc(nb ? nb : 1);
	memcpy(e->hdrs, g_hdrs, sizeof(vdl2_frame_hdr_t) * (size_t) nf);
	memcpy(e->bytes, g_bytes, nb);
	pthread_mutex_lock(&q_mtx);
	if (!q_started) {
		pthread_t th;
		pthread_create(&th, NULL, out_thread, NULL);
		q_started = 1;
	}
	if (q_tail)
		q_tail->next = e;
	else
		q_head = e;
	q_tail = e;
	pthread_cond_signal(&q_cond);
	pthread_mutex_unlock(&q_mtx);
}

void vdl2shim_finish(void)
{				/* what the last rounds completed is still on the device; then let the consumer print everything */
	pthread_mutex_lock(&g_busy);
	deliver();
	pthread_mutex_unlock(&g_busy);
	out_drain_wait();
}
#else
static void deliver(void)
{				/* called with g_busy held, after a process call returned */
	vdl2_block_t *out = g_blocks;
	int n = 0;
	if (vdl2_drain_blocks(g_gpu, out, SHIM_QCAP, &n))
		die("vdl2_drain_blocks");
	/* the library hands blocks over oldest trigger first; the reference queues a block when its burst ENDS (decodeVdlm2 at the
	   end of GETFEC, d8psk.c:199-204), so re-order by completion before they reach the single consumer */
	static int order[SHIM_QCAP];
	for (int i = 0; i < n; i++) {
		int j = i - 1;
		while (j >= 0 && (out[order[j]].end_dump > out[i].end_dump || (out[order[j]].end_dump == out[i].end_dump && out[order[j]].chn > out[i].chn))) {
			order[j + 1] = order[j];
			j--;
		}
		order[j + 1] = i;
	}
	for (int k = 0; k < n; k++) {
		const int i = order[k];
		channel_t *ch = NULL;
		for (int c = 0; c < nbch; c++)
			if (g_ch[c] && g_ch[c]->chn == out[i].chn)
				ch = g_ch[c];
		if (!ch)
			continue;
		msgblk_t *blk = ch->blk;
		stamp(&blk->tv, out[i].sync_dump);
		blk->ppm = out[i].ppm;
		blk->nbrow = out[i].nbrow;
		blk->nlbyte = out[i].nlbyte;
		for (int r = 0; r < 8; r++)
			memcpy(blk->data[r], out[i].data[r], 255);
		decodeVdlm2(ch);	/* takes blk, installs a fresh zeroed one (vdlm2.c:189-205) */
	}
	g_last_n = n;
}

void vdl2shim_finish(void)
{
	pthread_mutex_lock(&g_busy);	/* what the last rounds completed is still on the device */
	deliver();
	pthread_mutex_unlock(&g_busy);
	/* The reference's stopVdlm2() (vdlm2.c:182-187) waits for its queue to be EMPTY, not for the block blk_thread
	   has already taken off it, and main() exits right after (main.c:244-246).  A dongle delivers blocks over time;
	   a replay can hand over its whole last batch at the very end, so give the consumer time for it (about 50 us per
	   block on a current core) before the caller goes on to stopVdlm2(). */
	usleep(100000 + 500 * (unsigned)g_last_n);
}
#endif

/* One round of the reference's barrier protocol (or one batch of a replay).  The samples are copied into a page-locked ring
   slot of the handle and the upload + kernel are only ENQUEUED (vdl2_submit_copy): the caller goes straight back to its
   barrier while the GPU works, which is what makes the drop-in faster than the reference inside the reference's own
   protocol (the synchronous version spent 240 us per 32768-sample round in front of a 20 us kernel).  Completed blocks are
   collected -- with a synchronising drain -- only when the counter mirror says some are waiting, i.e. one round late and a
   few times per second per channel; vdl2shim_finish() collects the rest. */
void vdl2shim_feed(const void *iq, size_t nsamples)
{
	pthread_mutex_lock(&g_busy);
	if (g_first < 0)
		g_first = now();
	g_fed += nsamples;
	if (vdl2_submit_copy(g_gpu, iq, nsamples, 0))
		die("vdl2_submit_copy");
	int pending = 0;
	if (vdl2_pending_blocks(g_gpu, &pending))
		die("vdl2_pending_blocks");
	if (pending > 0 || getenv("VDL2_SHIM_SYNC"))
		deliver();
	pthread_mutex_unlock(&g_busy);
}

void vdl2shim_feed_rtl(const void *cu8, size_t nsamples)
{
	pthread_mutex_lock(&g_busy);
	if (vdl2_process_host_rtl(g_gpu, cu8, nsamples))
		die("vdl2_process_host_rtl");
	deliver();
	pthread_mutex_unlock(&g_busy);
}

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
#ifdef VDL2_SHIM_FILE
	pthread_barrier_wait(&Bar2);	/* parked for good: runFileSample() (file_shim.c) does all the work and main() exits */
#else
	const int leader = (param->chn == 0);
	if (leader)
#ifdef WITH_AIR
		vdl2shim_open(SDRINRATE, SDRCLK, VDL2_FMT_F32REAL, RTLINBUFSZ / 2);
#else
		vdl2shim_open(SDRINRATE, SDRCLK, VDL2_FMT_CF32, RTLINBUFSZ / 2);
#endif
	do {
		pthread_barrier_wait(&Bar2);
		if (leader)
			vdl2shim_feed(Cbuff, RTLINBUFSZ / 2);
		pthread_barrier_wait(&Bar1);
	} while (1);
#endif
	return NULL;
}
