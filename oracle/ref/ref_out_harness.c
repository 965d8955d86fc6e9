/* TEST INFRASTRUCTURE ONLY: the reference's output stage compiled IN PLACE (out.c, outacars.c, outxid.c, label.c, cJSON.c,
   crc.c are on the link line of oracle/Makefile; nothing of them is copied here) behind one call that returns the JSON line
   out() prints for a frame.  Used to pin oracle/port/vdl2_avlc_port.c (row f4).  The globals are the ones main.c defines
   (main.c:36-49); reversebits() is d8psk.c's (d8psk.c:39-52), restated because d8psk.c drags the whole demodulator in. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vdlm2.h"
#include "../orc_avlc_api.h"

int verbose = 0;
int grndmess = 1, emptymess = 1, undecmess = 0;
int jsonout = 1, routeout = 0, regout = 0;
char *netOutJsonAddr = NULL, *netOutSbsAddr = NULL;
char *idstation = "";
FILE *logfd;

unsigned int reversebits(const unsigned int bits, const int n)
{
	unsigned int r = 0;
	for (int i = 0; i < n; i++)
		if (bits & (1u << i))
			r |= 1u << (n - 1 - i);
	return r;
}

static int run_out(const uint8_t * hdata, int l, int chn, int Fr, float ppm, double t, char *buf, int cap);

int orc_out_json(const uint8_t * hdata, int l, int chn, int Fr, float ppm, double t, char *buf, int cap)
{
	verbose = 0;
	jsonout = 1;
	undecmess = 0;
	return run_out(hdata, l, chn, Fr, ppm, t, buf, cap);
}

/* the text out() prints at the default verbosity with -G -E -U (header line, "Command/Response from <addr> (...) to <addr>",
   link control line, payload): verbose 1 never reaches the hex dumps whose buffer the reference overruns */
int orc_out_text(const uint8_t * hdata, int l, int chn, int Fr, float ppm, double t, char *buf, int cap)
{
	verbose = 1;
	jsonout = 0;
	undecmess = 1;
	return run_out(hdata, l, chn, Fr, ppm, t, buf, cap);
}

static int run_out(const uint8_t * hdata, int l, int chn, int Fr, float ppm, double t, char *buf, int cap)
{
	static unsigned char copy[65 * 249];	/* out() strips the ACARS parity bits in place (outacars.c:225) */
	char *mem = NULL;
	size_t len = 0;
	msgblk_t blk;
	memset(&blk, 0, sizeof blk);
	blk.chn = chn;
	blk.Fr = Fr;
	blk.ppm = ppm;
	blk.tv.tv_sec = (time_t) t;
	blk.tv.tv_usec = (suseconds_t) ((t - (double)blk.tv.tv_sec) * 1e6);
	memset(copy, 0, sizeof copy);
	memcpy(copy, hdata, (size_t) l);
	logfd = open_memstream(&mem, &len);
	out(&blk, copy, l);
	fclose(logfd);
	if ((int)len >= cap)
		len = (size_t) cap - 1;
	memcpy(buf, mem, len);
	buf[len] = 0;
	free(mem);
	return (int)len;
}

/* seconds for `reps` passes of the reference's out() (JSON mode, -J -G -E) over n frames given as packed bytes:
   frame i = bytes[off[i] .. off[i] + len[i]).  The CPU side of the row-f4 measurement (bench.py `avlc.cpu_baseline`):
   it includes the reference's formatting, which the device record deliberately leaves to the host -- said so in the bench. */
#include <time.h>
double orc_out_time(const uint8_t * bytes, const uint32_t * off, const int32_t * len, int n, int reps)
{
	static char buf[60000];
	struct timespec a, b;
	clock_gettime(CLOCK_MONOTONIC, &a);
	for (int r = 0; r < reps; r++)
		for (int i = 0; i < n; i++)
			orc_out_json(bytes + off[i], len[i], 0, 136975000, 0.0f, 0.0, buf, (int)sizeof buf);
	clock_gettime(CLOCK_MONOTONIC, &b);
	return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
