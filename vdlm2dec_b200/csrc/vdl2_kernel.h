/* vdl2_kernel.h -- internal interface between vdl2_host.cu and vdl2_kernel.cu */
#ifndef VDL2_KERNEL_H
#define VDL2_KERNEL_H
#include "../../include/vdl2gpu.h"
#include "vdl2_common.h"

#ifndef VDL2_NSTAGE
#define VDL2_NSTAGE 3
#endif
#ifndef VDL2_MIN_CTAS
#define VDL2_MIN_CTAS 16
#endif		/* TMA boxes (32 rows x 128 B) in flight per warp */

#ifndef VDL2_D8_NST
#define VDL2_D8_NST 3		/* per-dump TMA boxes (32 rows x 64 B) in flight per warp, integer mixer */
#endif

#ifndef VDL2_MM_NST
#define VDL2_MM_NST 4		/* 64-byte column boxes (32 rows x 64 B) in the ring per warp, tensor-core mixer */
#endif

#ifdef __cplusplus
extern "C" {
#endif
int vdl2_kernel_smem_bytes(int nco_entries, int dp4a);
int vdl2_kernel_launch(int fmt, int dp4a, const void *tmap, const Vdl2KParams * kp, int grid, int smem, void *stream, int pdl);
int vdl2_kernel_occupancy(int fmt, int dp4a, int smem, int *ctas_per_sm);
int vdl2_kernel_upload_tables(const struct Vdl2Tables *t);
int vdl2_kernel_nsmid(unsigned *d_scratch_word, unsigned *out);
int vdl2_kernel_upload_sched(int slot, const unsigned *sched);
/* row f3 (vdl2_channelise_kernel): one pass over every stream -> the decimated streams of all its channels */
int vdl2_channelise_launch(int fmt, const void *tmap, int nstreams, int cps, int nrows, int nbox, int sched_slot, const void *bt, const void *dt,
			   void *out, size_t out_pitch, unsigned *ticket, int grid, void *stream);
#ifdef __cplusplus
}
#endif
#endif
