/*
 * file_shim.c -- raw-capture replay front end of the drop-in.
 *
 * The reference DECLARES a file front end but never defines or calls it:
 *     extern int initFile(char *file);        vdlm2.h:110
 *     extern int runFileSample(void);         vdlm2.h:111
 * This object defines both.  To be usable from the reference's UNMODIFIED main.c it also takes the place of
 * rtl.o: it exports what rtl.c exports -- initRtl / runRtlSample (vdlm2.h:102-103), SDRINRATE, SDRCLK
 * (rtl.c:36-37) and Fc (rtl.c:39) -- so that
 *     vdlm2dec [-v -J ...] -r capture.cu8 136.975 136.875
 * reads the capture named after -r instead of opening dongle number N.  initRtl parses the frequency list and
 * chooses the centre frequency by the rule of rtl.c:123-160 / 216-246, then calls initFile(); runRtlSample is
 * runFileSample.  It is linked next to d8psk_shim.o built with -DVDL2_SHIM_FILE (and, for the block pipeline
 * on the device, -DVDL2_SHIM_LINK); no librtlsdr, no Cbuff, no per-block barriers:
 *
 *   reference (rtl.c:273-295)                          here
 *   one 32768-sample callback per barrier round        batches of 2^22 samples (VDL2_FILE_BATCH) per launch
 *   u8 -> complex float on the host, 8 B/sample        raw bytes to the GPU (2 B/sample), converted in the kernel
 *   pageable Cbuff                                     ring of two page-locked buffers (vdl2_host_alloc), a reader
 *                                                      thread fills one while the other is being demodulated
 *
 * Capture format: by file extension -- .cu8 (default) .cs8 .cs16 .cf32 -- or VDL2_FILE_FORMAT; 2 Msps unless
 * VDL2_FILE_RATE says otherwise (SDRCLK = rate / 4000, as air.c:138 does for its rates).  With -v or
 * VDL2_FILE_STATS=1 the sample count and the steady-state rate of the replay are reported on stderr at the end.
 *
 * VDL2_RTL_QUIRK=1 (cu8 only) reproduces what the reference's callback does to the stream (rtl.c:285-292: the
 * index is incremented before the store, so slot 0 of every block keeps its zero and the last sample of the
 * block is dropped; partial reads are discarded, rtl.c:278-281).  The bytes still cross PCIe raw; the expansion
 * to the reference's complex-float block layout happens on the device (vdl2_process_host_rtl).  The output is
 * then identical to the reference fed the same bytes by a dongle (tests/test_replay.py).  VDL2_RTL_QUIRK=host
 * expands on the host instead, exactly like rtl.c (8 B/sample upload; kept as a cross-check).  Without the
 * variable every sample of the capture is demodulated once, in order.
 *
 * Airspy build (-DWITH_AIR, air.c instead of rtl.c): the object takes the place of air.o the same way -- initAirspy /
 * runAirspySample (vdlm2.h:106-107), SDRINRATE, SDRCLK, Fc (air.c:37-40).  main.c calls initAirspy(argv, optind, tparam)
 * after the options (main.c:205), so there is no argument left for a file name: the capture is named by VDL2_FILE.  It
 * holds float32 REAL samples (AIRSPY_SAMPLE_FLOAT32_REAL, air.c:123) at 6 Msps, or 5 Msps with VDL2_FILE_RATE=5000000
 * (the two rates air.c accepts, air.c:134-138); the centre follows air.c:48-70 and the mixer offsets are relative to
 * Fc + fs/4 (air.c:180-185).  rx_callback (air.c:190-217) copies samples unchanged, so there is no indexing mode here;
 * it only releases whole 32768-sample blocks, whereas the replay also demodulates a trailing partial block.
 *
 * Errors follow the reference (rtl.c:200-204, main.c:209-213): message on stderr, non-zero return from init,
 * exit(1) once running.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <semaphore.h>
#include <time.h>
#include "vdlm2.h"		/* the reference's header, found through -I */
#include "vdl2gpu.h"
#include "shim_internal.h"

extern int nbch;		/* main.c:59 */
extern int verbose;		/* main.c:36 */

#ifdef WITH_AIR
unsigned int SDRINRATE = 6000000;	/* air.c:37 */
unsigned int SDRCLK = 1500;	/* air.c:38 */
#define DEFAULT_FORMAT VDL2_FMT_F32REAL
#else
unsigned int SDRINRATE = 2000000;	/* rtl.c:36 */
unsigned int SDRCLK = 500;	/* rtl.c:37 */
#define DEFAULT_FORMAT VDL2_FMT_CU8
#endif
unsigned int Fc;		/* rtl.c:39, air.c:40 */

static FILE *g_file;
static int g_format = DEFAULT_FORMAT;
static int g_quirk;
static size_t g_batch = (size_t) 1 << 22;	/* samples per launch */

static size_t sample_bytes(int fmt)
{
	switch (fmt) {
	case VDL2_FMT_CS16:
	case VDL2_FMT_F32REAL:
		return 4;
	case VDL2_FMT_CF32:
		return 8;
	default:
		return 2;
	}
}

static int format_of(const char *name)
{
	if (!name)
		return -1;
	if (!strcasecmp(name, "cu8") || !strcasecmp(name, "u8") || !strcasecmp(name, "bin") || !strcasecmp(name, "raw"))
		return VDL2_FMT_CU8;
	if (!strcasecmp(name, "cs8") || !strcasecmp(name, "s8"))
		return VDL2_FMT_CS8;
	if (!strcasecmp(name, "cs16") || !strcasecmp(name, "s16"))
		return VDL2_FMT_CS16;
#ifdef WITH_AIR
	if (!strcasecmp(name, "f32real") || !strcasecmp(name, "f32") || !strcasecmp(name, "real"))
		return VDL2_FMT_F32REAL;
#endif
	if (!strcasecmp(name, "cf32") || !strcasecmp(name, "f32") || !strcasecmp(name, "cfile"))
		return VDL2_FMT_CF32;
	return -1;
}

int initFile(char *file)
{
	if (!file) {
		fprintf(stderr, "Need a capture file name\n");
		return 1;
	}
	g_file = strcmp(file, "-") ? fopen(file, "rb") : stdin;	/* "-": a pipe, e.g. from rtl_sdr */
	if (!g_file) {
		fprintf(stderr, "Failed to open capture %s\n", file);
		return 1;
	}
	const char *dot = strrchr(file, '.');
	int f = format_of(getenv("VDL2_FILE_FORMAT"));
	if (f < 0 && getenv("VDL2_FILE_FORMAT")) {
		fprintf(stderr, "Unknown capture format %s\n", getenv("VDL2_FILE_FORMAT"));
		return 1;
	}
	if (f < 0)
		f = format_of(dot ? dot + 1 : NULL);
	g_format = f < 0 ? DEFAULT_FORMAT : f;
#ifdef WITH_AIR
	if (g_format != VDL2_FMT_F32REAL) {	/* the channel offsets of this build are those of a real-sample stream (air.c:180-185) */
		fprintf(stderr, "The Airspy build replays float32 real captures only\n");
		return 1;
	}
#endif
	if (getenv("VDL2_FILE_RATE")) {
		const long r = atol(getenv("VDL2_FILE_RATE"));
		if (r < 1000000 || r % 4000) {
			fprintf(stderr, "Unusable capture rate %ld\n", r);
			return 1;
		}
		SDRINRATE = (unsigned)r;
		SDRCLK = SDRINRATE / 4000;	/* 21 / SDRCLK decimated samples per input sample at every rate (air.c:138) */
	}
	if (getenv("VDL2_FILE_BATCH") && atol(getenv("VDL2_FILE_BATCH")) > 0)
		g_batch = (size_t) atol(getenv("VDL2_FILE_BATCH"));
	const char *q = getenv("VDL2_RTL_QUIRK");
	g_quirk = 0;
	if (q && g_format == VDL2_FMT_CU8)
		g_quirk = !strcasecmp(q, "host") ? 2 : (atoi(q) ? 1 : 0);	/* 1: expanded on the device, 2: on the host */
	if (g_quirk)		/* whole callbacks only */
		g_batch = (g_batch + RTLINBUFSZ / 2 - 1) / (RTLINBUFSZ / 2) * (RTLINBUFSZ / 2);
	if (verbose > 1)
		fprintf(stderr, "Replaying %s: format %d, %u samples/s, %zu samples per launch%s\n", file, g_format, SDRINRATE, g_batch,
			g_quirk == 1 ? ", rtl.c block indexing (on the device)" : g_quirk ? ", rtl.c block indexing (on the host)" : "");
	return 0;
}

/* ---- ring of two page-locked buffers: the reader thread owns a slot between sem_free and sem_full ---- */
static struct slot {
	void *raw;		/* bytes as read from the capture */
	size_t nsamples;
} g_slot[2];
static sem_t g_free, g_full;

static void *reader(void *arg)
{
	const size_t bps = sample_bytes(g_format);
	(void)arg;
	for (int k = 0;; k ^= 1) {
		sem_wait(&g_free);
		size_t got = fread(g_slot[k].raw, 1, g_batch * bps, g_file);
		if (g_quirk)	/* a partial callback is discarded (rtl.c:278-281) */
			got -= got % RTLINBUFSZ;
		g_slot[k].nsamples = got / bps;
		sem_post(&g_full);
		if (g_slot[k].nsamples < g_batch)	/* end of the capture: the last slot is short (maybe empty) */
			return NULL;
	}
}

/* rtl.c:285-292 on one callback's worth of bytes: slot 0 untouched (zero), sample k in slot k + 1, last one lost */
static void expand_like_rtl(const unsigned char *in, float *out, size_t nsamples)
{
	const size_t blk = RTLINBUFSZ / 2;
	for (size_t b = 0; b + blk <= nsamples; b += blk) {
		const unsigned char *s = in + 2 * b;
		float *d = out + 2 * b;
		d[0] = 0.0f;
		d[1] = 0.0f;
		for (size_t k = 0; k + 1 < blk; k++) {
			d[2 * k + 2] = (float)s[2 * k] - (float)127.37;
			d[2 * k + 3] = (float)s[2 * k + 1] - (float)127.37;
		}
	}
}

int runFileSample(void)
{
	if (!g_file) {
		fprintf(stderr, "No capture opened\n");
		return 1;
	}
	pthread_barrier_wait(&Bar1);	/* every rcv_thread has registered its channel (d8psk_shim.c); they stay parked */
	vdl2shim_open(SDRINRATE, SDRCLK, g_quirk ? VDL2_FMT_CF32 : g_format, g_batch);

	const size_t bps = sample_bytes(g_format);
	float *wide = NULL;
	for (int k = 0; k < 2; k++)
		if (vdl2_host_alloc(g_batch * bps, &g_slot[k].raw)) {
			fprintf(stderr, "vdl2gpu replay: %s\n", vdl2_last_error(NULL));
			exit(1);
		}
	if (g_quirk == 2 && vdl2_host_alloc(g_batch * 8, (void **)&wide)) {
		fprintf(stderr, "vdl2gpu replay: %s\n", vdl2_last_error(NULL));
		exit(1);
	}
	sem_init(&g_free, 0, 2);
	sem_init(&g_full, 0, 0);
	pthread_t th;
	pthread_create(&th, NULL, reader, NULL);

	unsigned long long total = 0;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);	/* handle and ring exist: what follows is the steady state of a replay */
	for (int k = 0;; k ^= 1) {
		sem_wait(&g_full);
		const size_t n = g_slot[k].nsamples;
		if (n) {
			if (g_quirk == 1)
				vdl2shim_feed_rtl(g_slot[k].raw, n);
			else if (g_quirk) {
				expand_like_rtl(g_slot[k].raw, wide, n);
				vdl2shim_feed(wide, n);
			} else
				vdl2shim_feed(g_slot[k].raw, n);
			total += n;
		}
		sem_post(&g_free);
		if (n < g_batch)
			break;
	}
	pthread_join(th, NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	vdl2shim_finish();
	if (verbose > 1 || getenv("VDL2_FILE_STATS")) {
		const double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
		fprintf(stderr, "Replayed %llu samples in %.4f s (%.1f Msamples/s per channel, %d channels)\n", total, dt,
			dt > 0 ? 1e-6 * (double)total / dt : 0.0, vdl2shim_nch());
	}
	if (g_file != stdin)
		fclose(g_file);
	g_file = NULL;
	for (int k = 0; k < 2; k++)
		vdl2_host_free(g_slot[k].raw);
	vdl2_host_free(wide);
	return 0;
}

/* The frequency list both front ends of the reference accept (rtl.c:218-237, air.c:84-110): MHz as decimal text, at most
   MAXNBCHANNELS, values outside the aeronautical band skipped with a warning; channel numbers follow the order given.
   Fills param[].chn / .Fr and freq[] (Hz, same order), sets nbch; returns 0 or the reference's error (message printed). */
static int channel_list(char **argv, int first, thread_param_t * param, unsigned int *freq)
{
	nbch = 0;
	for (int i = first; argv[i] != NULL && nbch < MAXNBCHANNELS; i++) {
		const unsigned int hz = (int)(1000000 * atof(argv[i]));
		if (hz < 118000000 || hz > 138000000) {
			fprintf(stderr, "WARNING: Invalid frequency %d\n", hz);
			continue;
		}
		freq[nbch] = hz;
		param[nbch].chn = nbch;
		param[nbch].Fr = hz;
		nbch++;
	}
	if (nbch == 0) {
		fprintf(stderr, "Need a least one frequency\n");
		return 1;
	}
	return 0;
}

#ifdef WITH_AIR
/* ---- the air.o seam, so that the unmodified main.c drives the replay ---- */

/* Centre frequency, the rule of air.c:48-70.  At 5 Msps (Airspy R2) the reference narrows the R820T2 IF filter around the
   channels and shifts the tuning by the offset of the chosen pass band; the edge frequencies below are the tuner's
   (air.c:45-46).  There is no tuner to program here, but a capture taken by the reference was tuned this way. */
static const unsigned int if_hi[6] = { 1953050, 1980748, 2001344, 2032592, 2060291, 2087988 };
static const unsigned int if_lo[8] = { 525548, 656935, 795424, 898403, 1186034, 1502073, 1715133, 1853622 };

static unsigned int centre_for_real(unsigned int fmin, unsigned int fmax)
{
	const unsigned int need = fmax - fmin + 2 * STEPRATE;
	unsigned int shift = 0;
	if (SDRINRATE == 5000000) {
		int lo = 7, hi = 5;
		while (lo >= 0 && if_hi[5] - if_lo[lo] < need)	/* highest low edge that still leaves room */
			lo--;
		if (lo < 0)
			return 0;
		while (hi >= 0 && if_hi[hi] - if_lo[lo] > need)	/* lowest high edge that still leaves room */
			hi--;
		hi++;
		if (hi > 5)	/* exact fit of the widest pass band: the reference indexes past its table here */
			hi = 5;
		shift = (if_hi[hi] + if_lo[lo]) / 2 - SDRINRATE / 4;
	}
	return ((fmax + fmin) / 2 + shift + STEPRATE / 2) / STEPRATE * STEPRATE;
}

int initAirspy(char **argv, int optind, thread_param_t * param)
{				/* main.c:205, after the options: argv[optind...] are the frequencies (air.c:84-110) */
	unsigned int freq[MAXNBCHANNELS], fmin = 140000000, fmax = 0;
	if (channel_list(argv, optind, param, freq))
		return 1;
	for (int n = 0; n < nbch; n++) {
		if (freq[n] < fmin)
			fmin = freq[n];
		if (freq[n] > fmax)
			fmax = freq[n];
	}
	if (!getenv("VDL2_FILE")) {
		fprintf(stderr, "Name the capture to replay in VDL2_FILE\n");
		return 1;
	}
	if (initFile(getenv("VDL2_FILE")))
		return 1;
	if (SDRINRATE != 5000000 && SDRINRATE != 6000000) {	/* air.c:134-146 */
		fprintf(stderr, "did not find needed sampling rate\n");
		return -1;
	}
	if (getenv("VDL2_FILE_FC"))
		Fc = (unsigned int)(1000000 * atof(getenv("VDL2_FILE_FC")));
	else
		Fc = centre_for_real(fmin, fmax);
	if (Fc == 0) {
		fprintf(stderr, "Frequencies too far apart\n");
		return 1;
	}
	if (verbose > 1)
		fprintf(stderr, "Set freq. to %d hz\n", Fc);
	const unsigned int f0 = Fc + SDRINRATE / 4;	/* air.c:180-185 */
	for (int n = 0; n < nbch; n++)
		param[n].Fo = param[n].Fr - f0;
	return 0;
}

int runAirspySample(void)
{
	return runFileSample();
}
#else
/* ---- the rtl.o seam, so that the unmodified main.c drives the replay ---- */

/* Centre frequency for a set of channels, the rule of rtl.c:123-160: sorted ascending; from 50 kHz above the
   highest downwards (1 Hz steps) the first value that keeps every channel at least 2 STEPRATE away from the
   centre and from the band edge and is not the midpoint of two neighbours. */
static unsigned int centre_for(unsigned int *fd, int n)
{
	for (int i = 1; i < n; i++)	/* insertion sort, ascending */
		for (int j = i; j > 0 && fd[j - 1] > fd[j]; j--) {
			const unsigned int t = fd[j];
			fd[j] = fd[j - 1];
			fd[j - 1] = t;
		}
	const int guard = 2 * STEPRATE, half = (int)(SDRINRATE / 2);
	if (fd[n - 1] - fd[0] > SDRINRATE - 2 * guard) {
		fprintf(stderr, "Frequencies too far apart\n");
		return 0;
	}
	int fc = (int)fd[n - 1] + guard;
	for (; fc > (int)fd[0] - guard; fc--) {
		int ok = 1;
		for (int i = 0; i < n && ok; i++) {
			const int d = abs(fc - (int)fd[i]);
			if (d > half - guard || d < guard)
				ok = 0;
			else if (i > 0 && fc - (int)fd[i - 1] == (int)fd[i] - fc)
				ok = 0;
		}
		if (ok)
			break;
	}
	return (unsigned int)fc;
}

int initRtl(char **argv, int optind, thread_param_t * param)
{				/* main.c:141-143 calls this at "-r": argv[optind] is the capture, the rest the frequencies */
	unsigned int fd[MAXNBCHANNELS];
	if (argv[optind] == NULL) {
		fprintf(stderr, "Need a capture file name after -r\n");
		exit(1);
	}
	char *file = argv[optind];
	if (channel_list(argv, optind + 1, param, fd))
		return 1;
	if (initFile(file))	/* before the centre: VDL2_FILE_RATE changes the usable span */
		return 1;
	if (getenv("VDL2_FILE_FC"))	/* the capture was taken at a known centre */
		Fc = (unsigned int)(1000000 * atof(getenv("VDL2_FILE_FC")));
	else
		Fc = centre_for(fd, nbch);
	if (Fc == 0)
		return 1;
	for (int n = 0; n < nbch; n++)
		param[n].Fo = param[n].Fr - Fc;
	if (verbose > 1)
		fprintf(stderr, "Set center freq. to %dHz\n", (int)Fc);
	return 0;
}

int runRtlSample(void)
{
	return runFileSample();
}
#endif
