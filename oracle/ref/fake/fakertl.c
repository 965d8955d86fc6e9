/* TEST INFRASTRUCTURE: a file-backed librtlsdr so the UNMODIFIED rtl.c + main.c of the reference
   link and run without hardware.  VDL2_FAKE_IQ names a cu8 capture;
   rtlsdr_read_async() replays it in buf_len-byte callbacks and returns at end of file. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rtl-sdr.h"
struct rtlsdr_dev { int x; };
static struct rtlsdr_dev g_dev;
uint32_t rtlsdr_get_device_count(void) { return 1; }
int rtlsdr_get_device_usb_strings(uint32_t i, char *m, char *p, char *s) { strcpy(m, "fake"); strcpy(p, "file"); strcpy(s, "0"); return 0; }
const char *rtlsdr_get_device_name(uint32_t i) { return "fake rtl (file replay)"; }
int rtlsdr_open(rtlsdr_dev_t ** dev, uint32_t index) { *dev = &g_dev; return 0; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t * d, int m) { return 0; }
int rtlsdr_set_tuner_gain(rtlsdr_dev_t * d, int g) { return 0; }
int rtlsdr_get_tuner_gains(rtlsdr_dev_t * d, int *gains) { if (gains) gains[0] = 450; return 1; }
int rtlsdr_set_freq_correction(rtlsdr_dev_t * d, int ppm) { return 0; }
int rtlsdr_set_center_freq(rtlsdr_dev_t * d, uint32_t f) { fprintf(stderr, "fakertl: Fc=%u\n", f); return 0; }
int rtlsdr_set_sample_rate(rtlsdr_dev_t * d, uint32_t r) { return 0; }
int rtlsdr_reset_buffer(rtlsdr_dev_t * d) { return 0; }
int rtlsdr_read_async(rtlsdr_dev_t * d, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t n, uint32_t len)
{
	const char *path = getenv("VDL2_FAKE_IQ");
	FILE *f = path ? fopen(path, "rb") : NULL;
	if (!f) { fprintf(stderr, "fakertl: set VDL2_FAKE_IQ to a cu8 file\n"); return -1; }
	unsigned char *buf = malloc(len);
	while (fread(buf, 1, len, f) == len)
		cb(buf, len, ctx);
	free(buf);
	fclose(f);
	return 0;
}
