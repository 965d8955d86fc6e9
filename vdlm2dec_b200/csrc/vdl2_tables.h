/*
 * vdl2_tables.h -- constant tables of the VDL Mode 2 D8PSK front end, as DATA / formulas.
 *
 * Plain C, no code from the reference.  What each table is and where the reference keeps it:
 *   SYNC_K[17]   : unique-word phases in units of pi/8; the reference's SW[l] is
 *                  (float)(k * M_PI / 8) evaluated in double (d8psk.h:20-26).
 *   RC_HALF[32]  : rising half of the 63-tap raised-cosine (alpha 0.6) interpolating
 *                  low-pass, 16 taps per zero crossing, peak 1.0 at tap 31; the reference
 *                  stores 63 ten-digit literals for MFLTLEN 65 slots, so taps 63 and 64
 *                  are implicit zeros (d8psk.h:28-45, vdlm2.h:37).  The literals are not
 *                  bit-reproducible from the closed form, so they
 *                  are carried as numeric data.
 *   soft demap   : P(bit=1 | phase index i), i in [0,256], von Mises kappa=10 posterior
 *                  over the 8 sector centres (2j+1)pi/8, printed with "%f" (six decimals);
 *                  bit 1 (MSB) <- sectors {-1,-3,-5,-7}, bit 2 <- {+-5,+-7}, bit 3 <- {+-3,+-5}
 *                  (ggrey.c:60-103, d8psk.h:47-249).  vdl2_make_softmap() regenerates them.
 */
#ifndef VDL2_TABLES_H
#define VDL2_TABLES_H

#define VDL2_NBPH 17		/* symbols in the sync fit (vdlm2.h:54) */
#define VDL2_D8DWN 4		/* WSYNC steps per symbol (vdlm2.h:55) */
#define VDL2_PHRING (VDL2_NBPH * VDL2_D8DWN)
#define VDL2_MFLTLEN 65		/* vdlm2.h:37 */
#define VDL2_MBUFLEN 17		/* vdlm2.h:38 */
#define VDL2_STEPRATE 25000	/* vdlm2.h:33 */

static const int VDL2_SYNC_K[VDL2_NBPH] = {
	2, 3, 10, 15, 8, 9, 12, 9, 2, 5, 4, 9, 4, 1, -4, -5, 2
};

/* taps 0..31 (tap 31 is the unit centre); tap 62-j == tap j */
static const double VDL2_RC_HALF[32] = {
	-0.0063474526, -0.0147744088, -0.0251715417, -0.0372531112,
	-0.0505438764, -0.0643762574, -0.0778990609, -0.0900984580,
	-0.0998311862, -0.1058691815, -0.1069540690, -0.1018592183,
	-0.0894564364, -0.0687838818, -0.0391114778, +0.0000000000,
	+0.0486498533, +0.1065617468, +0.1730641128, +0.2470886715,
	+0.3271881497, +0.4115732615, +0.4981679546, +0.5846808858,
	+0.6686901328, +0.7477373336, +0.8194268281, +0.8815249907,
	+0.9320548266, +0.9693810568, +0.9922813460, +1.0000000000
};

#ifndef __CUDACC_RTC__
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

/* sync[l] = (float)(k_l * pi / 8), product and quotient in double like the C literals */
static inline void vdl2_make_sync(float sync[VDL2_NBPH])
{
	for (int l = 0; l < VDL2_NBPH; l++)
		sync[l] = (float)(VDL2_SYNC_K[l] * 3.14159265358979323846 / 8);
}

static inline void vdl2_make_mflt(float mflt[VDL2_MFLTLEN])
{
	for (int i = 0; i < VDL2_MFLTLEN; i++) {
		if (i < 32)
			mflt[i] = (float)VDL2_RC_HALF[i];
		else if (i < 63)
			mflt[i] = (float)VDL2_RC_HALF[62 - i];
		else
			mflt[i] = 0.0f;
	}
}

/* soft[b][i], b = 0..2 (MSB first), i = 0..256 */
static inline void vdl2_make_softmap(float soft[3][257])
{
	const double PI = 3.14159265358979323846;
	const double kappa = 10.0;
	/* membership of the sector centre (2j+1)pi/8, j = -4..3, in "bit = 1" */
	for (int b = 0; b < 3; b++) {
		for (int i = -128; i <= 128; i++) {
			double p1 = 0, pt = 0;
			static const int centre[8] = { 1, 3, 5, 7, -1, -3, -5, -7 };	/* ggrey.c order */
			for (int q = 0; q < 8; q++) {
				int c = centre[q];
				double pb = exp(kappa * cos(c * PI / 8 - i * PI / 128));
				int a = c < 0 ? -c : c;
				int one = (b == 0) ? (c < 0) : (b == 1) ? (a == 5 || a == 7) : (a == 3 || a == 5);
				pt += pb;
				if (one)
					p1 += pb;
			}
			char txt[32];
			snprintf(txt, sizeof txt, "%f", p1 / pt);
			soft[b][i + 128] = (float)strtod(txt, NULL);
		}
	}
}
#endif
#endif
