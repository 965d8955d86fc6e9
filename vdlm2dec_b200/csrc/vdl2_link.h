/* vdl2_link.h -- internal interface between vdl2_host.cu and vdl2_link.cu (block pipeline on the GPU) */
#ifndef VDL2_LINK_H
#define VDL2_LINK_H
#include <stdint.h>
#include "vdl2_common.h"

/* identical layouts to vdl2_frame_t / vdl2_blkstat_t (include/vdl2gpu.h) */
struct Vdl2FrameRec {
	int32_t block, len, chn, Fr;
	float ppm;
	int32_t pad;		/* end_dump - sync_dump of the block (vdl2_frame_t.pad) */
	int64_t sync_dump;
	uint8_t hdata[2016];
};
struct Vdl2BlkStat {
	int8_t rs[8];
	int32_t nbytes, nframes;
};

#ifdef __cplusplus
extern "C" {
#endif
int vdl2_link_upload_tables(void);
int vdl2_link_launch(const Vdl2BlockRec * d_blocks, int nblocks, Vdl2FrameRec * d_frames, unsigned *d_nframes, unsigned cap,
		     Vdl2BlkStat * d_stats, uint8_t * d_rows_after, void *stream);
/* vdl2_avlc.cu (row f4): one 48-byte field record per frame */
int vdl2_avlc_launch(const Vdl2FrameRec * d_frames, int nframes, void *d_recs, void *stream);
/* frames (unordered, count on the device) -> completion-order rank, 16-byte aligned offsets, 32-byte headers + packed bytes,
   field records in the same order; d_totals[0] = frames, [1] = bytes.  expect = host-side upper bound of the frame count */
int vdl2_frames_pack_launch(const Vdl2FrameRec * d_frames, const unsigned *d_nframes, unsigned cap, int *d_rank, unsigned *d_offs,
			    unsigned *d_totals, void *d_hdrs, uint8_t * d_bytes, unsigned bytes_cap, void *d_recs, int expect, void *d_keys, void *stream);
#ifdef __cplusplus
}
#endif
#endif
