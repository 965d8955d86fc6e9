/*
 * oracle/port/vdl2_link_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Independent plain-C restatement of the block pipeline behind the demodulator
 * (reference: blk_thread, vdlm2.c:84-161):
 *   per row   errors-and-erasures decoding of the shortened RS(255,249) code over GF(2^8)
 *             (rs.c:81-291: field polynomial 0x187, first consecutive root alpha^120, 6 roots);
 *             the last row erases the check bytes that were never transmitted (set_eras, vdlm2.c:63-82);
 *   stream    HDLC bit un-stuffing across the rows, LSB first (vdlm2.c:116-147);
 *   framing   flag 0x7e delimits; every closing flag triggers check_frame (vdlm2.c:40-61): at least 13
 *             bytes and PPP FCS16 residue 0xf0b8 over the bytes between the flags (crc.c, crc.h:3).
 * Written from the algorithm, with polynomial-form field arithmetic and generated tables (the reference
 * uses index form and literal tables), and pinned against the reference compiled in place
 * (oracle/_ref/libvdl2linkref.so) by tests/test_link_oracle.py.  The CUDA kernel mirrors THIS file.
 *
 * Behaviours of the reference that are reproduced on purpose:
 *   - rs()'s result is ignored: an uncorrectable row passes through as received;
 *   - hdata[0] is never cleared while waiting for the first flag: bytes are OR-ed into it until the
 *     accumulated value equals 0x7e (vdlm2.c:119-139 with k == 0);
 *   - the byte index is not reset after a frame: later frames of the same block start at hdata[1]
 *     too, so only the first frame of a block can pass the FCS.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../orc_link_api.h"

/* ------------------------------------------------------------------ GF(2^8) */
static uint8_t gf_exp[512], gf_log[256];
static uint16_t fcs_tab[256];
static int tables_ready;

static void make_tables(void)
{
	if (tables_ready)
		return;
	int x = 1;
	for (int i = 0; i < 255; i++) {
		gf_exp[i] = gf_exp[i + 255] = (uint8_t) x;
		gf_log[x] = (uint8_t) i;
		x <<= 1;
		if (x & 0x100)
			x ^= 0x187;	/* x^8 + x^7 + x^2 + x + 1 (rs.c:17-49 is its antilog table) */
	}
	gf_log[0] = 255;
	for (int b = 0; b < 256; b++) {	/* reflected CCITT polynomial 0x8408 (crc.c) */
		unsigned c = (unsigned)b;
		for (int k = 0; k < 8; k++)
			c = (c & 1u) ? (c >> 1) ^ 0x8408u : c >> 1;
		fcs_tab[b] = (uint16_t) c;
	}
	tables_ready = 1;
}

static inline uint8_t gmul(uint8_t a, uint8_t b)
{
	return (a && b) ? gf_exp[gf_log[a] + gf_log[b]] : 0;
}

static inline uint8_t gpow(int e)
{				/* alpha^e, any integer e */
	e %= 255;
	return gf_exp[e < 0 ? e + 255 : e];
}

static inline uint8_t gdiv(uint8_t a, uint8_t b)
{				/* b != 0 */
	return a ? gf_exp[gf_log[a] + 255 - gf_log[b]] : 0;
}

#define NR 6			/* check symbols */
#define FCR 120			/* first consecutive root */

/* rs.c:81-291.  Returns the number of corrected symbols, -1 if uncorrectable.  data[0] is the highest-degree
   coefficient: position p carries x^(254-p). */
int port_rs(uint8_t data[255], const int *eras, int neras)
{
	uint8_t S[NR], lam[NR + 1], B[NR + 1], T[NR + 1], om[NR];
	int nz = 0;
	for (int i = 0; i < NR; i++) {	/* S_i = data(alpha^(FCR+i)) by Horner */
		const uint8_t a = gpow(FCR + i);
		uint8_t s = 0;
		for (int j = 0; j < 255; j++)
			s = gmul(s, a) ^ data[j];
		S[i] = s;
		nz |= s;
	}
	if (!nz)
		return 0;
	memset(lam, 0, sizeof lam);
	lam[0] = 1;
	for (int e = 0; e < neras; e++) {	/* erasure locator: prod (1 + x alpha^(254 - pos)) */
		const uint8_t X = gpow(254 - eras[e]);
		for (int j = e + 1; j > 0; j--)
			lam[j] ^= gmul(lam[j - 1], X);
	}
	memcpy(B, lam, sizeof B);
	int L = neras;
	for (int r = neras + 1; r <= NR; r++) {	/* Berlekamp-Massey, rs.c:136-187 */
		uint8_t d = 0;
		for (int i = 0; i < r; i++)
			d ^= gmul(lam[i], S[r - i - 1]);
		if (d == 0) {
			memmove(B + 1, B, NR);
			B[0] = 0;
			continue;
		}
		T[0] = lam[0];
		for (int i = 0; i < NR; i++)
			T[i + 1] = lam[i + 1] ^ gmul(d, B[i]);
		if (2 * L <= r + neras - 1) {
			L = r + neras - L;
			for (int i = 0; i <= NR; i++)
				B[i] = gdiv(lam[i], d);
		} else {
			memmove(B + 1, B, NR);
			B[0] = 0;
		}
		memcpy(lam, T, sizeof lam);
	}
	int deg = 0;
	for (int i = 0; i <= NR; i++)
		if (lam[i])
			deg = i;
	/* Chien search: lambda(alpha^i) = 0 <=> an error at position i - 1 (rs.c:198-222) */
	int root[NR], loc[NR], count = 0;
	for (int i = 1; i <= 255 && count < deg; i++) {
		uint8_t q = 1;
		for (int j = 1; j <= deg; j++)
			if (lam[j])
				q ^= gmul(lam[j], gpow(i * j));
		if (q == 0) {
			root[count] = i;
			loc[count] = i - 1;
			count++;
		}
	}
	if (count != deg)
		return -1;
	int dom = 0;
	for (int i = 0; i < NR; i++) {	/* omega = S * lambda mod x^NR (rs.c:232-245) */
		uint8_t t = 0;
		for (int j = (deg < i ? deg : i); j >= 0; j--)
			t ^= gmul(S[i - j], lam[j]);
		om[i] = t;
		if (t)
			dom = i;
	}
	for (int j = count - 1; j >= 0; j--) {	/* Forney (rs.c:251-283), last root first like the reference */
		uint8_t num = 0, den = 0;
		for (int i = dom; i >= 0; i--)
			num ^= gmul(om[i], gpow(i * root[j]));
		const int top = (deg < NR - 1 ? deg : NR - 1) & ~1;
		for (int i = top; i >= 0; i -= 2)
			den ^= gmul(lam[i + 1], gpow(i * root[j]));
		if (den == 0)
			return -1;	/* corrections already applied stay, like in the reference */
		if (num)
			data[loc[j]] ^= gdiv(gmul(num, gpow(root[j] * (FCR - 1))), den);
	}
	return count;
}

/* set_eras, vdlm2.c:63-82 */
static int last_row_erasures(int nlbyte, int eras[4])
{
	if (nlbyte <= 30) {
		eras[0] = 251; eras[1] = 252; eras[2] = 253; eras[3] = 254;
		return 4;
	}
	if (nlbyte <= 67) {
		eras[0] = 253; eras[1] = 254;
		return 2;
	}
	return 0;
}

typedef void (*frame_cb) (void *ctx, const uint8_t * hdata, int l);

/* One block.  rows[8][255] is corrected in place.  Returns k, the index of the hdata byte under construction at the end. */
static int link_block(uint8_t rows[8][255], int nbrow, int nlbyte, int8_t rs_out[8], frame_cb cb, void *ctx, int *nframes)
{
	static __thread uint8_t hd[8 * 249 + 8];
	int k = 0, nbits = 0, ones = 0;
	unsigned cur = 0;	/* hdata[k] under construction */
	memset(hd, 0, sizeof hd);
	for (int r = 0; r < nbrow && r < 8; r++) {
		int eras[4], ne = 0, by = 249;
		if (r == nbrow - 1) {
			by = nlbyte;
			ne = last_row_erasures(nlbyte, eras);
		}
		const int c = port_rs(rows[r], eras, ne);
		if (rs_out)
			rs_out[r] = (int8_t) c;
		for (int i = 0; i < by; i++) {
			for (int n = 0; n < 8; n++) {
				if ((rows[r][i] >> n) & 1) {
					cur |= 1u << nbits;
					ones++;
				} else {
					const int stuffed = (ones == 5);	/* exactly five ones before: a stuffing zero */
					ones = 0;
					if (stuffed)
						continue;
				}
				if (++nbits < 8)
					continue;
				nbits = 0;
				hd[k] = (uint8_t) cur;
				if (cur == 0x7e) {
					if (k == 0) {
						k = 1;
						cur = 0;
					} else if (k == 1) {
						cur = 0;	/* flag right after the opening flag: still waiting for data */
					} else {
						const int l = k + 1;
						if (l >= 13) {	/* check_frame, vdlm2.c:40-61 */
							unsigned crc = 0xffff;
							for (int q = 1; q < l - 1; q++)
								crc = (crc >> 8) ^ fcs_tab[(crc ^ hd[q]) & 0xff];
							if (crc == 0xf0b8) {
								if (nframes)
									(*nframes)++;
								if (cb)
									cb(ctx, hd, l);
							}
						}
						k++;
						cur = 0;
					}
				} else if (k > 0) {
					k++;
					cur = 0;
				}	/* k == 0 and not a flag: keep OR-ing into hdata[0] */
			}
		}
	}
	return k;
}

struct cap {
	orc_frame *frames;
	int max, n, block, overflow;
	const orc_block *src;
};

static void capture(void *ctx, const uint8_t * hdata, int l)
{
	struct cap *c = (struct cap *)ctx;
	if (!c->frames)
		return;
	if (c->n >= c->max) {
		c->overflow = 1;
		return;
	}
	orc_frame *f = &c->frames[c->n++];
	memset(f, 0, sizeof *f);
	f->block = c->block;
	f->len = l;
	f->chn = c->src->chn;
	f->Fr = c->src->Fr;
	f->ppm = c->src->ppm;
	f->sync_dump = c->src->sync_dump;
	memcpy(f->hdata, hdata, l < ORC_FRAME_MAX ? l : ORC_FRAME_MAX);
}

int orc_link_decode(const orc_block * blocks, int n, orc_frame * frames, int max_frames, int *n_frames, orc_blkstat * stats,
		    uint8_t * rows_after)
{
	make_tables();
	struct cap c = { frames, max_frames, 0, 0, 0, NULL };
	for (int i = 0; i < n; i++) {
		uint8_t rows[8][255];
		memcpy(rows, blocks[i].data, sizeof rows);
		orc_blkstat st;
		memset(&st, 0, sizeof st);
		int nf = 0;
		c.block = i;
		c.src = &blocks[i];
		st.nbytes = link_block(rows, blocks[i].nbrow, blocks[i].nlbyte, st.rs, capture, &c, &nf);
		st.nframes = nf;
		if (stats)
			stats[i] = st;
		if (rows_after)
			memcpy(rows_after + (size_t) i * 8 * 255, rows, sizeof rows);
	}
	if (n_frames)
		*n_frames = c.n;
	return c.overflow;
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
	return "port";
}
