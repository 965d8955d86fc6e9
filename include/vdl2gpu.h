/*
 * vdl2gpu.h -- C ABI of the B200-native VDL Mode 2 front-end DSP path.
 *
 * Drop-in scope: exactly the per-IQ-sample hot path of TLeconte/vdlm2dec
 *   sample conversion            rtl.c:285-292 / air.c:206-208
 *   NCO mix + integrate-and-dump d8psk.c:343-382 (rcv_thread)
 *   interpolating filter, sync fit, timing, D8PSK slicer   d8psk.c:219-333
 *   soft demap, descrambler, header decode, de-interleave  d8psk.c:54-217, viterbi.c:37-96
 * up to the hand-off of a completed msgblk_t (decodeVdlm2(), vdlm2.c:189), and -- optionally, the next row of
 * the scope table -- the block pipeline behind it (blk_thread, vdlm2.c:84-161: rs.c, HDLC, crc.c) up to the
 * call of out().  Everything downstream (out*.c, label.c, cJSON) stays host code of the reference.
 *
 * The reference has no plugin API: the seam is the object d8psk.o (vdlm2.h:113-114,128).
 * Two layers are exported by libvdl2gpu.so:
 *   1. the batch API below (plain pointers and sizes), which is what a cgo/ctypes/FFI
 *      binding or the reference's C code calls;
 *   2. libvdl2shim.so / d8psk_gpu.o, which re-exports the reference
 *      symbols rcv_thread / initD8psk / reversebits on top of layer 1.
 *
 * Conventions follow the reference (rtl.c:200-204): int return, 0 = OK, non-zero = failure
 * with a message retrievable through vdl2_last_error() (and printed on stderr).
 * There is NO CPU fallback: every entry point fails if no sm_100 device is usable.
 */
#ifndef VDL2GPU_H
#define VDL2GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDL2_ABI_VERSION 4

/* input sample formats; CU8 is what rtl.c receives (rtl.c:287-289: x - 127.37f),
   CF32 is the already converted Cbuff of the reference (vdlm2.h:89),
   F32REAL is the Airspy real-sample mode (air.c:123,206-208). */
enum vdl2_format {
	VDL2_FMT_CU8 = 0,	/* unsigned 8-bit I,Q; value - 127.37f */
	VDL2_FMT_CS8 = 1,	/* signed 8-bit I,Q */
	VDL2_FMT_CS16 = 2,	/* signed 16-bit I,Q */
	VDL2_FMT_CF32 = 3,	/* float32 I,Q */
	VDL2_FMT_F32REAL = 4	/* float32 real samples */
};

/* optional observation points; they cost HBM writes, keep off in production */
#define VDL2_TAP_DUMPS 1u	/* T1: the 84 ksps decimated stream */
#define VDL2_TAP_STEPS 2u	/* T2: P, err, fr of every idle (WSYNC) step */
#define VDL2_TAP_SYNCS 4u	/* T3: trigger events */
#define VDL2_TAP_SYMS  8u	/* T4/T5: per-symbol differential phase, Gray index, 3 soft bits */
/* option bit in the same mask: evaluate the full 17-point sync fit at EVERY idle step like the
   reference does (d8psk.c:259-289) instead of only where the screen says err < 4 is possible;
   same outputs, slower; implied by VDL2_TAP_STEPS */
#define VDL2_OPT_EXACT_IDLE 0x100u
/* option bit: mix cu8/cs8 input with the generic fp32 mixer (the one every other format uses) instead of
   the integer dot-product mixer that is the default for 8-bit input at 2 Msps; same outputs within the
   parity tolerance, slower; for A/B tests */
#define VDL2_OPT_FLOAT_MIX 0x200u
/* option bit: mix cu8/cs8 input at 2 Msps with the IDP.4A integer mixer of round 1 instead of the int8 tensor-core
   mixer that is the default since ABI version 4 (same exact integer sums, slower; for A/B tests) */
#define VDL2_OPT_DP4A_MIX 0x800u
/* option bit: let consecutive vdl2_process_device() launches overlap (programmatic dependent launch: the next launch's
   warps move in while the tail of the previous one drains; the per-channel order is kept by the kernel's own flags).
   No per-launch events are recorded then: vdl2_stats_t.last_kernel_ms stays 0 -- bracket with your own events on
   vdl2_cuda_stream().  Results are identical. */
#define VDL2_OPT_OVERLAP 0x400u

/* mirrors thread_param_t (vdlm2.h:49-52) */
typedef struct {
	int chn;		/* channel number reported in blocks */
	int Fr;			/* channel frequency, Hz */
	int Fo;			/* offset from the tuner centre, Hz (multiple of 25 kHz, d8psk.c:348-357) */
} vdl2_chan_param_t;

typedef struct {
	unsigned fs;		/* SDRINRATE (rtl.c:36 / air.c:37) */
	unsigned sdrclk;	/* SDRCLK    (rtl.c:37 / air.c:38,138) */
	int format;		/* enum vdl2_format */
	int nch;		/* channels */
	int ch_per_stream;	/* channels demodulated from each input stream (1..8); nch % ch_per_stream == 0 */
	int device;		/* CUDA device ordinal */
	unsigned taps;		/* VDL2_TAP_* mask */
	size_t max_samples;	/* largest nsamples (per stream) of one vdl2_process_* call */
	int max_blocks;		/* capacity of the completed-block queue between drains (0 = default) */
} vdl2_config_t;

/* completed block = the fields of msgblk_t the demodulator owns (vdlm2.h:39-47):
   ppm (d8psk.c:302), nbrow/nlbyte (d8psk.c:94-95), data[r][0..254] (d8psk.c:127,176).
   tv is wall-clock in the reference (d8psk.c:295); here the trigger position is given
   in 84 kHz dump units and in input samples so the caller can synthesise it. */
typedef struct {
	int64_t sync_dump;	/* index (since create) of the decimated sample that triggered */
	int64_t end_dump;	/* index of the decimated sample that completed the block */
	int32_t chn;
	int32_t Fr;
	float ppm;
	int32_t nbrow;
	int32_t nlbyte;
	uint8_t data[8][255];
	uint8_t pad[4];
} vdl2_block_t;			/* 2080 bytes */

typedef struct {
	int64_t dump;
	float P, err, fr;
	int32_t pad;
} vdl2_step_t;

typedef struct {
	int64_t dump;
	int32_t clk;
	float df, ppm, P1;
} vdl2_sync_t;

typedef struct {
	int64_t dump;
	float D, P;
	int32_t gi;
	float v[3];
	int32_t state_after;
	int32_t pad;
} vdl2_sym_t;

typedef struct {
	uint64_t kernel_launches;	/* front-end kernel launches since create */
	uint64_t samples_in;		/* per-stream samples accepted */
	uint64_t samples_done;		/* per-stream samples demodulated (whole 1 ms rows) */
	uint64_t blocks_out;		/* blocks completed */
	uint64_t blocks_dropped;	/* blocks lost to a full queue */
	float last_kernel_ms;		/* device time of the last launch (CUDA events) */
	int n_sm;
	int grid;			/* persistent warps launched */
	int smem_bytes;			/* dynamic shared memory per warp-CTA */
	float last_link_ms;		/* device time of the last block-pipeline launch */
	uint32_t link_launches;
	uint64_t frames_out;
} vdl2_stats_t;

typedef struct vdl2gpu vdl2gpu_t;

int vdl2_abi_version(void);
const char *vdl2_last_error(const vdl2gpu_t * h);	/* h may be NULL: error of the last failed create */

int vdl2_create(const vdl2_config_t * cfg, const vdl2_chan_param_t * chans, vdl2gpu_t ** out);
int vdl2_destroy(vdl2gpu_t * h);

/* Host input: iq holds nstreams = nch/ch_per_stream streams, stream s at iq + s*pitch_bytes,
   nsamples samples each (interleaved I,Q in cfg.format).  Copies H2D, demodulates every
   complete 1 ms row (the sub-millisecond tail is kept for the next call, like the
   reference carries clk/nf/no across blocks, d8psk.c:343-347) and returns when done. */
int vdl2_process_host(vdl2gpu_t * h, const void *iq, size_t nsamples, size_t pitch_bytes);

/* The asynchronous half of vdl2_process_host(): enqueues the upload and the launch and returns; iq must stay valid (and
   should be page-locked, vdl2_host_alloc) until vdl2_sync() or a drain.  vdl2_process_host() = vdl2_submit_host() + vdl2_sync(). */
int vdl2_submit_host(vdl2gpu_t * h, const void *iq, size_t nsamples, size_t pitch_bytes);
/* Same, but the samples are first copied into a page-locked ring slot owned by the handle: iq may be reused as soon as the
   call returns.  This is what the drop-in for the reference's barrier protocol uses (Cbuff is overwritten by the next SDR
   callback, rtl.c:283-294): the 32768-sample round costs one 256 KB memcpy and three enqueues on the host instead of a
   synchronous pageable upload, a kernel wait and a drain. */
int vdl2_submit_copy(vdl2gpu_t * h, const void *iq, size_t nsamples, size_t pitch_bytes);
/* Number of completed blocks waiting in the queue, as of the newest launch that has FINISHED (never waits; may lag behind
   by the launch in flight).  Lets a streaming caller skip the (synchronising) drain while nothing is pending. */
int vdl2_pending_blocks(vdl2gpu_t * h, int *n_out);

/* Device input, same layout, already resident in HBM.  Zero-copy when no tail is pending
   and nsamples is a whole number of rows; asynchronous on the handle's stream. */
int vdl2_process_device(vdl2gpu_t * h, const void *d_iq, size_t nsamples, size_t pitch_bytes);
int vdl2_sync(vdl2gpu_t * h);

/* ---- wideband shared-stream channeliser (generalises d8psk.c:353-381, where every channel
   thread re-reads the whole Cbuff): ONE pass over each input stream writes the decimated 84 ksps streams (d8psk.c:374-381)
   of all the channels of the handle that listen to it: d_out[ch * out_pitch + row * 84 + k] = dump k of millisecond `row` of
   channel ch, as interleaved (re, im) floats; out_pitch counts complex values (even, >= rows * 84).  Device pointers, same
   input layout as vdl2_process_device(), whole 1 ms rows, 8-bit input at 2 Msps.  No demodulator state is touched;
   asynchronous on the handle's stream (vdl2_sync); vdl2_stats_t.last_kernel_ms reports its device time.  Every output equals
   the fused kernel's own decimated stream (tap VDL2_TAP_DUMPS) bit for bit. ---- */
int vdl2_channelise_device(vdl2gpu_t * h, const void *d_iq, size_t nsamples, size_t pitch_bytes, float *d_out, size_t out_pitch);

/* completed blocks since the last drain, oldest trigger first (the msgblk_t hand-off) */
int vdl2_drain_blocks(vdl2gpu_t * h, vdl2_block_t * out, int max, int *n_out);

/* taps (only what cfg.taps enabled); each read returns and clears the channel's records */
int vdl2_read_dumps(vdl2gpu_t * h, int ch, float *iq_out, size_t max, size_t *n_out);
int vdl2_read_steps(vdl2gpu_t * h, int ch, vdl2_step_t * out, size_t max, size_t *n_out);
int vdl2_read_syncs(vdl2gpu_t * h, int ch, vdl2_sync_t * out, size_t max, size_t *n_out);
int vdl2_read_syms(vdl2gpu_t * h, int ch, vdl2_sym_t * out, size_t max, size_t *n_out);

/* ---- block pipeline behind the demodulator (blk_thread, vdlm2.c:84-161):
   per row rs() (rs.c:81-291), HDLC bit un-stuffing, flag framing, FCS16 check (check_frame, vdlm2.c:40-61).
   A frame is exactly what the reference passes to out(blk, hdata, l) (vdlm2.h:134). ---- */
typedef struct {
	int32_t block;		/* index into the blocks of the same call */
	int32_t len;		/* l: bytes in hdata including both flags */
	int32_t chn, Fr;	/* copied from the block */
	float ppm;
	int32_t pad;		/* end_dump - sync_dump of the block: length of the burst in 84 kHz dumps */
	int64_t sync_dump;
	uint8_t hdata[2016];
} vdl2_frame_t;			/* 2048 bytes */

typedef struct {
	int8_t rs[8];		/* rs() result per row: symbols corrected, -1 uncorrectable (the reference ignores it) */
	int32_t nbytes;		/* un-stuffed bytes consumed into hdata[] at the end of the block */
	int32_t nframes;	/* frames that passed check_frame() */
} vdl2_blkstat_t;

/* blocks from host memory -> frames, in (block, position) order.  stats (nblocks) and rows_after
   (nblocks * 8 * 255 bytes: data[][] after the rs() calls) may be NULL. */
int vdl2_link_decode(vdl2gpu_t * h, const vdl2_block_t * blocks, int nblocks, vdl2_frame_t * frames, int max_frames, int *n_frames,
		     vdl2_blkstat_t * stats, uint8_t * rows_after);
/* like vdl2_drain_blocks, but the completed blocks go through the block pipeline ON THE DEVICE first (no
   round trip): returns the frames (oldest trigger first, then channel, then position) and, if blocks != NULL,
   the blocks themselves in the same order (frame.block indexes them; -1 when the blocks were not asked for) */
int vdl2_drain_frames(vdl2gpu_t * h, vdl2_frame_t * frames, int max_frames, int *n_frames, vdl2_block_t * blocks, int max_blocks,
		      int *n_blocks);

/* ---- ingest: page-locked host memory for the buffers handed to vdl2_process_host().
   From such a buffer the upload runs at the PCIe rate without the driver's staging copy; a replay front end
   (file_shim.c: initFile / runFileSample, vdlm2.h:110-111) keeps a ring of them so that reading the capture
   overlaps the demodulation of the previous batch.  Needs a usable device like every other entry point. ---- */
int vdl2_host_alloc(size_t bytes, void **out);
int vdl2_host_free(void *p);
/* Raw cu8 as an RTL dongle delivers it, demodulated the way the REFERENCE sees it: in_callback (rtl.c:285-292) increments
   its index before the store, so slot 0 of every 32768-sample block keeps its zero, sample k lands in slot k + 1 and the
   last sample of the block is lost.  The bytes cross PCIe raw (2 B/sample) and are expanded to that complex-float block
   layout ON THE DEVICE (a zero sample cannot be written in cu8: the conversion is u - 127.37), then demodulated like
   vdl2_process_host() input.  The handle must have format VDL2_FMT_CF32 and one stream; nsamples must be a whole number
   of 32768-sample callbacks.  Output is bit-identical to the reference fed the same bytes. */
int vdl2_process_host_rtl(vdl2gpu_t * h, const void *cu8, size_t nsamples);

/* ---- frame fields: what the reference's out() (out.c:517-570) and outacars()
   (outacars.c:214-290) derive from the bytes of a frame BEFORE they format text or JSON -- addresses through icaoaddr()
   (out.c:426-435), direction, command/response, on-ground bit, link control byte, payload class, and for ACARS the CRC
   verdict, the parity-stripped header fields and the extent of the message text.  Formatting stays on the host. ---- */
enum vdl2_avlc_kind {
	VDL2_AVLC_EMPTY = 0,		/* l <= 13: no information field (out.c:530) */
	VDL2_AVLC_XID = 1,		/* hdata[10] == 0x82 (out.c:562) */
	VDL2_AVLC_ACARS = 2,		/* ff ff 01 header, ACARS CRC good (out.c:566, outacars.c:222-230) */
	VDL2_AVLC_ACARS_BADCRC = 3,
	VDL2_AVLC_OTHER = 4		/* anything else with l > 13 ("unknown data", out.c:570) */
};

typedef struct {
	uint32_t faddr, taddr;	/* icaoaddr(&hdata[5]), icaoaddr(&hdata[1]): 3-bit type << 24 | 24-bit address */
	uint8_t fromair;	/* (faddr >> 24) == 1 */
	uint8_t rep;		/* 1 = response, 0 = command */
	uint8_t gnd;		/* aircraft on ground */
	uint8_t lc;		/* link control byte hdata[9] (outlinkctrl, out.c:484-504) */
	uint8_t kind;		/* enum vdl2_avlc_kind */
	uint8_t mode, ack, bid, bs, be;	/* ACARS, parity stripped; ack 0x15 -> '!', bid 0 -> ' ' (outacars.c:243-261) */
	uint8_t label[2];	/* label[1] 0x7f -> 'd' */
	uint8_t reg[7];		/* registration as sent (fixreg()'s formatting stays on the host) */
	uint8_t nno, nfid;	/* characters in no[] / fid[] */
	uint8_t no[4], fid[6];
	uint8_t pad;
	uint16_t txt_off, txt_len;	/* message text = hdata[txt_off .. txt_off + txt_len), each byte & 0x7f */
	uint16_t info_off, info_len;	/* information field of any kind: hdata[10 .. l - 3) */
} vdl2_avlc_t;			/* 48 bytes */

/* frames (as returned by vdl2_link_decode / vdl2_drain_frames) -> one record per frame, same order */
int vdl2_avlc_extract(vdl2gpu_t * h, const vdl2_frame_t * frames, int nframes, vdl2_avlc_t * recs);

/* ---- rows f1 + f4 end to end.  Like vdl2_drain_frames(), but the frames leave the device ORDERED and PACKED: frame i is
   hdrs[i] + bytes[hdrs[i].offset .. + hdrs[i].len) (exactly the (hdata, l) of out(), vdlm2.h:134; offsets are 16-byte
   aligned) and, if recs != NULL, its field record recs[i] -- computed on the device before the frame left HBM.  Order =
   completion order, the order in which the reference's single consumer sees the blocks (decodeVdlm2 at the end of a burst,
   vdlm2.c:189-206): end of the burst (sync_dump + dur), then channel, then position inside the block.  Buffers should come
   from vdl2_host_alloc() (page-locked); sizes: max_frames headers / records, max_bytes of frame bytes. ---- */
typedef struct {
	int64_t sync_dump;	/* trigger of the burst the frame came in (84 kHz dump index since create) */
	int32_t chn, Fr;
	float ppm;
	int32_t len;		/* l: bytes including both flags */
	uint32_t offset;	/* of the frame's first byte in `bytes` */
	int32_t dur;		/* end_dump - sync_dump */
} vdl2_frame_hdr_t;		/* 32 bytes */
int vdl2_drain_frames_packed(vdl2gpu_t * h, vdl2_frame_hdr_t * hdrs, int max_frames, int *n_frames, uint8_t * bytes, size_t max_bytes,
			     size_t *n_bytes, vdl2_avlc_t * recs);
/* device time (CUDA events) of the ranking, packing and field kernels of the last vdl2_drain_frames_packed() */
float vdl2_last_pack_ms(const vdl2gpu_t * h);

/* ---- several GPUs behind one handle: input stream s, with its channels, lives on
   devices[s mod ndev]; there is no exchange step, so there is no collective -- every device runs the same kernel on its own
   streams and the host merges the completed blocks into the order one device would have produced (oldest trigger first,
   then channel), which is what the single consumer of the reference expects (blk_thread's queue, vdlm2.c:189-206).
   cfg->device is ignored; cfg->nch / chans / the input layout are those of the WHOLE job, exactly as for vdl2_create() /
   vdl2_process_host().  The same ordinal may appear more than once (two handles on one GPU). ---- */
typedef struct vdl2multi vdl2multi_t;
int vdl2_multi_create(const vdl2_config_t * cfg, const vdl2_chan_param_t * chans, const int *devices, int ndev, vdl2multi_t ** out);
int vdl2_multi_destroy(vdl2multi_t * m);
/* uploads and launches on every device first, then waits for all of them */
int vdl2_multi_process_host(vdl2multi_t * m, const void *iq, size_t nsamples, size_t pitch_bytes);
int vdl2_multi_drain_blocks(vdl2multi_t * m, vdl2_block_t * out, int max, int *n_out);
int vdl2_multi_ndev(const vdl2multi_t * m);		/* devices that own at least one stream */
vdl2gpu_t *vdl2_multi_handle(vdl2multi_t * m, int i);	/* the per-device handle (statistics, taps) */
const char *vdl2_multi_last_error(const vdl2multi_t * m);	/* m may be NULL: error of the last failed create */

int vdl2_get_stats(vdl2gpu_t * h, vdl2_stats_t * st);
/* the CUDA stream the kernels run on (a cudaStream_t), so callers can bracket with events */
void *vdl2_cuda_stream(vdl2gpu_t * h);

#ifdef __cplusplus
}
#endif
#endif
