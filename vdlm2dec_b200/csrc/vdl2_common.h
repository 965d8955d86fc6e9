/*
 * vdl2_common.h -- structures shared by the host side of libvdl2gpu and the sm_100a kernel.
 * Vocabulary follows the reference: a "dump" is one 84 ksps integrate-and-dump output
 * (d8psk.c:374-381), a "step" is one idle-mode sync evaluation (d8psk.c:248-313),
 * a "row" is 1 ms of input (the joint period of the 21/SDRCLK dump clock and the 25 kHz
 * NCO table): fs/1000 samples, always 84 dumps.
 */
#ifndef VDL2_COMMON_H
#define VDL2_COMMON_H
#include <stdint.h>
#include <stddef.h>
#include <vector_types.h>

#define VDL2_DUMPS_PER_ROW 84
#define VDL2_ROWS_PER_TILE 32	/* one row per lane */
#define VDL2_TILE_DUMPS (VDL2_DUMPS_PER_ROW * VDL2_ROWS_PER_TILE)	/* 2688 */
#define VDL2_HIST 16		/* dumps of history in front of a tile (MBUFLEN-1) */
#define VDL2_PHHIST 64		/* idle-mode phases of history ((NBPH-1)*D8DWN) */
#define VDL2_MAX_CHUNKS 21000	/* 16-byte chunks per row: 336000 B (cf32 @ 42 Msps, the longest window of the config-5 sweep: 500 samples per dump) / 16 */
#define VDL2_SCHED_SLOTS 16	/* distinct (fs, SDRCLK, format) combinations alive in one process */
#define VDL2_W8_PHASES 104	/* integer mixer: oscillator entries by NCO phase, 80 + 24 so that a dump never wraps */
#define VDL2_W8_ENTRIES 120	/* ... plus 16 "first sample only" entries closing the 23-sample dumps of a row */
#define VDL2_MM_PHASES 10	/* int8 tensor-core mixer: window phases = NCO period / 8 samples (80 / 8 at 2 Msps) */
#define VDL2_MM_BT_ENTRIES (VDL2_MM_PHASES * 24)	/* uint4 per (phase, column 0..5, lane & 3): B fragments */
#define VDL2_MM_DT_ENTRIES ((VDL2_DUMPS_PER_ROW + 1) * 4)	/* int4 per (dump, lane & 3): accumulator start, scale, offset correction; +1: prefetch */
#define VDL2_MM_W 0x1u		/* schedule word bits of the tensor-core mixer, see vdl2_mma_tables.h */
#define VDL2_MM_R 0x2u
#define VDL2_SCR_WORDS 512	/* descrambler sequence: 25 + 8*8*255 = 16345 bits max */

#define VDL2_FLAG_NO_SCREEN 1u	/* debug: run the exact 17-point fit at every idle step */
#define VDL2_FLAG_NO_PREPASS 2u	/* A/B: no speculative pass A while waiting for the previous tile */
#define VDL2_TAP_DUMPS_BIT 1u
#define VDL2_TAP_STEPS_BIT 2u
#define VDL2_TAP_SYNCS_BIT 4u
#define VDL2_TAP_SYMS_BIT 8u

enum { VDL2_ST_WSYNC = 0, VDL2_ST_GETHEAD = 1, VDL2_ST_GETDATA = 2 };	/* GETFEC is folded into GETDATA */

/* per-channel demodulator state kept in HBM between tiles and calls: the GPU analogue of
   channel_t (vdlm2.h:56-79) plus the rcv_thread locals that survive a block (d8psk.c:343-347) */
struct Vdl2ChanState {
	float hist_re[VDL2_HIST], hist_im[VDL2_HIST];	/* last 16 dumps (Inbuff ring, oldest first) */
	float ph[VDL2_PHHIST];	/* last 64 idle-mode phases, ph[63] newest (Ph ring in logical order) */
	float hv[28];		/* descrambled header soft bits collected so far */
	float perr, p2err, pfr, df, P1, ppm;
	int32_t clk, state;
	int32_t symidx;		/* symbols sliced since the trigger */
	int32_t nbrow, nlbyte;	/* header values (d8psk.c:94-95) */
	int32_t bytes_done;	/* data+FEC bytes stored so far */
	int32_t bitacc, nbitacc;	/* partial byte, LSB first */
	int64_t sync_dump;
	int32_t chn, Fr;
	uint32_t n_steps, n_syncs, n_syms, n_dumps;	/* tap record counts */
	/* forecast for the speculative idle search of later tiles: as far as is known the channel is idle from global dump
	   fc_dump on, with tick clock fc_clk there.  Written at the end of an idle tile and, early, as soon as the header
	   of a burst gives its length; read without synchronisation (a wrong guess is detected, see idle_run) */
	int32_t pad[2];		/* the forecast below is 16-byte aligned: a waiting tile polls it with ONE vector load */
	int64_t fc_dump;
	int32_t fc_clk;
	float fc_df;		/* frequency offset of the burst that ends at fc_dump (BurstPre) */
};
static_assert(offsetof(Vdl2ChanState, fc_dump) % 16 == 0 && sizeof(Vdl2ChanState) % 16 == 0, "forecast: one aligned 16-byte load");

/* constant tables of the kernel (independent of rate and format, so handles can share them) */
struct Vdl2Tables {
	float mflt[68];		/* interpolating low-pass taps, 65 used (d8psk.h:28-45) */
	float sync[20];		/* unique-word phases, 17 used (d8psk.h:20-26) */
	float soft[3][260];	/* soft demap, 257 used per bit (d8psk.h:47-249) */
	unsigned scr[VDL2_SCR_WORDS];	/* descrambler bit sequence from seed 0x4D4B (d8psk.c:54-65,299) */
	unsigned sched_slots[VDL2_SCHED_SLOTS][VDL2_DUMPS_PER_ROW];	/* mixer dump schedules, one slot per distinct
									   (fs, SDRCLK, format) in use; see Vdl2KParams.sched_slot */
	unsigned char hcol[32];	/* header code parity-check columns, 25 used (viterbi.c:29-35) */
};

/* completed-block record: identical layout to vdl2_block_t (include/vdl2gpu.h) */
struct Vdl2BlockRec {
	int64_t sync_dump, end_dump;
	int32_t chn, Fr;
	float ppm;
	int32_t nbrow, nlbyte;
	uint8_t data[8 * 255];
	uint8_t pad[4];
};

struct Vdl2StepRec { int64_t dump; float P, err, fr; int32_t pad; };
struct Vdl2SyncRec { int64_t dump; int32_t clk; float df, ppm, P1; };
struct Vdl2SymRec  { int64_t dump; float D, P; int32_t gi; float v[3]; int32_t state_after; int32_t pad; };

/* kernel arguments (the TMA descriptor travels separately as a __grid_constant__) */
struct Vdl2KParams {
	int nch, ch_per_stream;
	int ntiles;		/* tiles of 32 rows in this launch */
	int nrows;		/* rows in this launch (last tile may be short) */
	int chunks_per_row;	/* row bytes / 16 */
	int nbox;		/* 128-byte column boxes per row */
	int nco_pairs;		/* NCO table length in float4 entries */
	int wext;		/* entries appended (copies of the head) so that a dump never wraps */
	int64_t dump_base;	/* global dump index of row 0 of this launch */
	Vdl2ChanState *state;
	const float4 *wtab;	/* [nch][nco_pairs]: (re[n], re[n+1], im[n], im[n+1]) */
	const float4 *dcorr;	/* [nch][84]: per dump (1/nf, 1/nf, -cre/nf, -cim/nf), see dump_close */
	const float *soft;	/* [3][260]: copy of Vdl2Tables.soft in global memory (per-lane indexed look-ups), may be NULL */
	const unsigned *scr;	/* [VDL2_SCR_WORDS]: copy of Vdl2Tables.scr, same reason (valid when soft is) */
	const uint4 *w8;	/* integer mixer only, [nch][VDL2_W8_ENTRIES]: (digit 2, digit 1, digit 0, 0) words of signed bytes
				   (wr[n], wi[n], wr[n+1], wi[n+1]); sched word = (first sample << 16) | (last entry << 8) | first entry */
	int sched_slot;		/* which c_tab.sched_slots[] row: per dump of a row, (w0 << 16) | (E << 8) | np: np whole 16-byte
				   chunks, then the chunk in which the dump ends after sample E; w0 = index of the dump's
				   first chunk in the (extended) oscillator table */
	unsigned *ticket;	/* work counters [0..2], rotating per launch (ticket_sel); [3] launches completed; [4] block queue count;
				   [5..7] CTAs exited per counter slot; [8] dropped; [12..15] statistics */
	int ticket_sel;
	int launch_seq;		/* launches of this handle before this one */
	int tile_base;		/* tiles per channel completed by earlier launches */
	int *progress;		/* [nch]: tiles completed since create */
	unsigned *slotmask;	/* [nsmid]: scratch slots in use per SM */
	int slots_per_sm;
	uint8_t *curblk;	/* [nch][2048] block under construction */
	float2 *scratch;	/* [grid][VDL2_HIST + VDL2_TILE_DUMPS]: decimated stream of the tile a warp works on */
	Vdl2BlockRec *outq;
	unsigned *outq_count;
	unsigned outq_cap;
	unsigned *dropped;
	unsigned taps;
	unsigned flags;		/* VDL2_FLAG_* */
	float2 *tap_dumps;	/* [nch][cap_dumps] */
	Vdl2StepRec *tap_steps;	/* [nch][cap_steps] */
	Vdl2SyncRec *tap_syncs;	/* [nch][cap_syncs] */
	Vdl2SymRec *tap_syms;	/* [nch][cap_syms] */
	unsigned cap_dumps, cap_steps, cap_syncs, cap_syms;
};

#endif
