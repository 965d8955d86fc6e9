/*
 * oracle/orc_link_api.h -- TEST INFRASTRUCTURE ONLY.
 * C API shared by the two CPU checkers of the block pipeline that follows the demodulator
 * (blk_thread, vdlm2.c:84-161 -- rs() per row, HDLC bit un-stuffing,
 * flag framing, FCS16 check, then out(blk, hdata, l)):
 *   oracle/ref/ref_link_harness.c  the reference's vdlm2.c + rs.c + crc.c compiled in place -> oracle/_ref/
 *   oracle/port/vdl2_link_port.c   an independent plain-C restatement -> oracle/libvdl2linkport.so
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load these libraries.
 */
#ifndef ORC_LINK_API_H
#define ORC_LINK_API_H
#include <stdint.h>
#include "orc_api.h"

#define ORC_FRAME_MAX 2016	/* >= 1 + 8*249 + 1: the longest hdata[] a block can produce */

typedef struct {
	int32_t block;		/* index of the block in the call */
	int32_t len;		/* l of out(blk, hdata, l): bytes including both flags */
	int32_t chn, Fr;
	float ppm;
	int32_t pad;
	int64_t sync_dump;
	uint8_t hdata[ORC_FRAME_MAX];
} orc_frame;			/* 32 + 2016 = 2048 B */

typedef struct {
	int8_t rs[8];		/* return value of rs() per row (corrected symbols, -1 = uncorrectable); rows >= nbrow: 0 */
	int32_t nbytes;		/* k at the end of the block: index of the hdata byte under construction */
	int32_t nframes;	/* frames that passed check_frame() */
} orc_blkstat;			/* 16 B */

#ifdef __cplusplus
extern "C" {
#endif
/* runs every block through the pipeline in order; frames in order of emission.
   rows_after (nullable): n * 8 * 255 bytes, data[][] after the rs() calls.  Returns 0, or 1 if max_frames was too small. */
int orc_link_decode(const orc_block * blocks, int n, orc_frame * frames, int max_frames, int *n_frames, orc_blkstat * stats,
		    uint8_t * rows_after);
/* seconds for `reps` passes over the n blocks (no capture) */
double orc_link_time(const orc_block * blocks, int n, int reps);
const char *orc_link_kind(void);
#ifdef __cplusplus
}
#endif
#endif
