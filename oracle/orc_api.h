/*
 * oracle/orc_api.h -- TEST INFRASTRUCTURE ONLY.
 * C API shared by the two CPU checkers of the hot path:
 *   oracle/ref/ref_harness.c  (the reference's own d8psk.c compiled in place -> oracle/_ref/)
 *   oracle/port/vdl2_port.c   (an independent plain-C restatement -> oracle/libvdl2port.so)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load these libraries.  The product (vdlm2dec_b200/) never does.
 *
 * Tap points follow SURVEY.md section 4.1:
 *   T1 dumps   : decimated 84 ksps stream D           (d8psk.c:377-378)
 *   T2 steps   : per WSYNC step P, err, fr            (d8psk.c:252,277-289)
 *   T3 syncs   : trigger events clk, df, ppm, P1      (d8psk.c:292-308)
 *   T4/T5 syms : per-symbol D, Grey index, soft bits  (d8psk.c:323-329,213-216)
 *   T6 blocks  : completed msgblk_t at decodeVdlm2()  (d8psk.c:201, vdlm2.c:189)
 */
#ifndef ORC_API_H
#define ORC_API_H
#include <stdint.h>
#include <stddef.h>

#define ORC_TAP_DUMPS  1u
#define ORC_TAP_STEPS  2u
#define ORC_TAP_SYNCS  4u
#define ORC_TAP_SYMS   8u
#define ORC_TAP_BLOCKS 16u

typedef struct {
	void *p;
	size_t n, cap;
} orc_vec;

typedef struct {
	int64_t dump;
	float P, err, fr;
	int32_t pad;
} orc_step;			/* 24 B */

typedef struct {
	int64_t dump;
	int32_t clk;
	float df, ppm, P1;
} orc_sync;			/* 24 B */

typedef struct {
	int64_t dump;
	float D, P;
	int32_t gi;
	float v[3];
	int32_t state_after;
	int32_t pad;
} orc_sym;			/* 40 B */

typedef struct {
	int64_t sync_dump, end_dump;
	int32_t chn, Fr;
	float ppm;
	int32_t nbrow, nlbyte;
	uint8_t data[8][255];
	uint8_t pad[4];
} orc_block;			/* 16+8+4+8+2040+4 = 2080 B */

#ifdef __cplusplus
extern "C" {
#endif
void *orc_open(int chn, int Fr, int Fo, unsigned fs, unsigned sdrclk, int real_input, uint32_t taps);
void orc_close(void *h);
void orc_feed_cf32(void *h, const float *iq, size_t n);
void orc_feed_f32real(void *h, const float *x, size_t n);
void orc_feed_cu8(void *h, const uint8_t * iq, size_t n, float offset);
void orc_feed_cs8(void *h, const int8_t * iq, size_t n);
void orc_feed_cs16(void *h, const int16_t * iq, size_t n);
void orc_feed_rtl_block_quirk(void *h, const uint8_t * blk);
const void *orc_tap(void *h, int which, size_t *count);
void orc_clear_taps(void *h);
int64_t orc_ndump(void *h);
const char *orc_kind(void);
const float *orc_table(int which);
/* oscillator table of an open handle (d8psk.c:353-357): up to max complex floats into out[2 * max]; returns the table length */
int orc_nco(void *h, float *out, int max);
double orc_time_cu8(int Fr, int Fo, unsigned fs, unsigned sdrclk, const uint8_t * iq, size_t n, int reps);
#ifdef __cplusplus
}
#endif
#endif
