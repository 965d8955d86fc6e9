/* TEST INFRASTRUCTURE: a file-backed libairspy so the UNMODIFIED air.c + main.c of the reference link and run without
   hardware.  VDL2_FAKE_IQ names a capture of float32 REAL samples (AIRSPY_SAMPLE_FLOAT32_REAL, air.c:123) at
   VDL2_FAKE_RATE samples/s (5000000 = Airspy R2, 6000000 = Mini; air.c:134-138).  airspy_start_rx() replays it from a
   thread in transfers of 49152 samples -- deliberately not a divisor of the reference's 32768-sample block, so that
   rx_callback's re-blocking (air.c:191-217) is exercised -- and airspy_is_streaming() turns false at end of file. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <libairspy/airspy.h>
struct airspy_device { int x; };
static struct airspy_device g_dev;
static volatile int g_streaming;
static airspy_sample_block_cb_fn g_cb;
static uint32_t fake_rate(void) { const char *r = getenv("VDL2_FAKE_RATE"); return r ? (uint32_t) atoi(r) : 6000000u; }
int airspy_open(struct airspy_device **d) { *d = &g_dev; return AIRSPY_SUCCESS; }
int airspy_open_sn(struct airspy_device **d, uint64_t sn) { (void)sn; *d = &g_dev; return AIRSPY_SUCCESS; }
int airspy_close(struct airspy_device *d) { (void)d; return AIRSPY_SUCCESS; }
int airspy_exit(void) { return AIRSPY_SUCCESS; }
const char *airspy_error_name(enum airspy_error e) { return e == AIRSPY_SUCCESS ? "AIRSPY_SUCCESS" : "AIRSPY_ERROR (fake)"; }
int airspy_set_sample_type(struct airspy_device *d, enum airspy_sample_type t) { (void)d; return t == AIRSPY_SAMPLE_FLOAT32_REAL ? AIRSPY_SUCCESS : AIRSPY_ERROR_OTHER; }
int airspy_get_samplerates(struct airspy_device *d, uint32_t * buf, const uint32_t len)
{
	(void)d;
	if (len == 0) { *buf = 2; return AIRSPY_SUCCESS; }	/* count query, air.c:130 */
	buf[0] = 10000000u;
	if (len > 1) buf[1] = fake_rate();
	return AIRSPY_SUCCESS;
}
int airspy_set_samplerate(struct airspy_device *d, uint32_t i) { (void)d; (void)i; return AIRSPY_SUCCESS; }
int airspy_set_packing(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return AIRSPY_SUCCESS; }
int airspy_set_linearity_gain(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return AIRSPY_SUCCESS; }
int airspy_set_freq(struct airspy_device *d, const uint32_t f) { (void)d; fprintf(stderr, "fakeairspy: Fc=%u\n", f); return AIRSPY_SUCCESS; }
int airspy_r820t_write(struct airspy_device *d, uint8_t r, uint8_t v) { (void)d; (void)r; (void)v; return AIRSPY_SUCCESS; }
static void *replay(void *arg)
{
	(void)arg;
	const char *path = getenv("VDL2_FAKE_IQ");
	FILE *f = path ? fopen(path, "rb") : NULL;
	if (!f) { fprintf(stderr, "fakeairspy: set VDL2_FAKE_IQ to a float32 real capture\n"); g_streaming = 0; return NULL; }
	enum { N = 49152 };
	float *buf = malloc(sizeof(float) * N);
	size_t got;
	while ((got = fread(buf, sizeof(float), N, f)) > 0) {
		airspy_transfer_t t;
		memset(&t, 0, sizeof t);
		t.device = &g_dev;
		t.samples = buf;
		t.sample_count = (int)got;
		t.sample_type = AIRSPY_SAMPLE_FLOAT32_REAL;
		g_cb(&t);
	}
	free(buf);
	fclose(f);
	g_streaming = 0;
	return NULL;
}
int airspy_start_rx(struct airspy_device *d, airspy_sample_block_cb_fn cb, void *ctx)
{
	(void)d; (void)ctx;
	pthread_t th;
	g_cb = cb;
	g_streaming = 1;
	return pthread_create(&th, NULL, replay, NULL) ? AIRSPY_ERROR_OTHER : AIRSPY_SUCCESS;
}
int airspy_is_streaming(struct airspy_device *d) { (void)d; return g_streaming ? AIRSPY_TRUE : 0; }
