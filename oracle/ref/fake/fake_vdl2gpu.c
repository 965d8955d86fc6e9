/* TEST INFRASTRUCTURE ONLY -- never part of the product, never linked into libvdl2gpu.so.
 *
 * A stand-in for the handful of libvdl2gpu entry points the drop-in shims call, answering from the CPU oracle
 * (oracle/libvdl2port.so).  Purpose: the HOST logic of the replay front end (vdlm2dec_b200/csrc/file_shim.c +
 * d8psk_shim.c -DVDL2_SHIM_FILE: argument seam, centre-frequency rule, reader ring, batching, the rtl.c index
 * quirk, hand-off to decodeVdlm2) can be checked against the all-reference binary in the CPU test tier, where
 * there is no GPU (tests/test_replay.py::test_replay_host_logic_*).  The binary built from it is
 * oracle/_ref/vdlm2dec_file_hostcheck; the product binaries link the real library and fail loudly without a
 * device. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vdl2gpu.h"
#include "../../orc_api.h"
#include "../../orc_link_api.h"

struct vdl2gpu {
	vdl2_config_t cfg;
	void *orc[8];
};

int vdl2_abi_version(void) { return VDL2_ABI_VERSION; }
const char *vdl2_last_error(const vdl2gpu_t * h) { (void)h; return "fake_vdl2gpu"; }

int vdl2_create(const vdl2_config_t * cfg, const vdl2_chan_param_t * chans, vdl2gpu_t ** out)
{
	if (cfg->nch > 8 || cfg->ch_per_stream != cfg->nch)
		return 1;
	fprintf(stderr, "fake_vdl2gpu: TEST STAND-IN answering from the CPU oracle -- not libvdl2gpu.so, not a product path\n");
	vdl2gpu_t *h = calloc(1, sizeof *h);
	h->cfg = *cfg;
	for (int c = 0; c < cfg->nch; c++)
		h->orc[c] = orc_open(chans[c].chn, chans[c].Fr, chans[c].Fo, cfg->fs, cfg->sdrclk, cfg->format == VDL2_FMT_F32REAL, ORC_TAP_BLOCKS);
	*out = h;
	return 0;
}

int vdl2_destroy(vdl2gpu_t * h)
{
	for (int c = 0; c < h->cfg.nch; c++)
		orc_close(h->orc[c]);
	free(h);
	return 0;
}

int vdl2_process_host(vdl2gpu_t * h, const void *iq, size_t n, size_t pitch)
{
	(void)pitch;
	if (n > h->cfg.max_samples)
		return 1;
	for (int c = 0; c < h->cfg.nch; c++)
		switch (h->cfg.format) {
		case VDL2_FMT_CU8: orc_feed_cu8(h->orc[c], iq, n, (float)127.37); break;
		case VDL2_FMT_CS8: orc_feed_cs8(h->orc[c], iq, n); break;
		case VDL2_FMT_CS16: orc_feed_cs16(h->orc[c], iq, n); break;
		case VDL2_FMT_CF32: orc_feed_cf32(h->orc[c], iq, n); break;
		default: orc_feed_f32real(h->orc[c], iq, n); break;
		}
	return 0;
}

int vdl2_submit_copy(vdl2gpu_t * h, const void *iq, size_t n, size_t pitch)
{				/* the stand-in has nothing asynchronous: the samples are consumed before the call returns */
	return vdl2_process_host(h, iq, n, pitch);
}

int vdl2_pending_blocks(vdl2gpu_t * h, int *n_out)
{
	size_t total = 0;
	for (int c = 0; c < h->cfg.nch; c++) {
		size_t k = 0;
		orc_tap(h->orc[c], ORC_TAP_BLOCKS, &k);
		total += k;
	}
	*n_out = (int)total;
	return 0;
}

int vdl2_process_host_rtl(vdl2gpu_t * h, const void *cu8, size_t n)
{				/* the oracle's own statement of rtl.c:285-292, one callback at a time */
	if (h->cfg.format != VDL2_FMT_CF32 || n % 32768 || n > h->cfg.max_samples)
		return 1;
	for (int c = 0; c < h->cfg.nch; c++)
		for (size_t b = 0; b < n; b += 32768)
			orc_feed_rtl_block_quirk(h->orc[c], (const uint8_t *)cu8 + 2 * b);
	return 0;
}

static int by_trigger(const void *a, const void *b)
{
	const vdl2_block_t *x = a, *y = b;
	if (x->sync_dump != y->sync_dump)
		return x->sync_dump < y->sync_dump ? -1 : 1;
	return x->chn - y->chn;
}

int vdl2_drain_blocks(vdl2gpu_t * h, vdl2_block_t * out, int max, int *n_out)
{
	int n = 0;
	for (int c = 0; c < h->cfg.nch; c++) {
		size_t k = 0;
		const orc_block *b = orc_tap(h->orc[c], ORC_TAP_BLOCKS, &k);
		if (n + (int)k > max)
			return 1;
		if (k)
			memcpy(out + n, b, k * sizeof *b);	/* orc_block and vdl2_block_t share one layout (2080 B) */
		n += (int)k;
		orc_clear_taps(h->orc[c]);
	}
	qsort(out, n, sizeof *out, by_trigger);
	*n_out = n;
	return 0;
}

int vdl2_drain_frames(vdl2gpu_t * h, vdl2_frame_t * frames, int max_frames, int *n_frames, vdl2_block_t * blocks, int max_blocks,
		      int *n_blocks)
{				/* the completed blocks, oldest trigger first, through the block pipeline oracle (vdlm2.c:84-161) */
	static vdl2_block_t own[4096];
	vdl2_block_t *b = blocks ? blocks : own;
	int nb = 0;
	*n_frames = 0;
	if (vdl2_drain_blocks(h, b, blocks ? max_blocks : 4096, &nb))
		return 1;
	if (n_blocks)
		*n_blocks = nb;
	if (nb == 0)
		return 0;
	if (orc_link_decode((const orc_block *)b, nb, (orc_frame *) frames, max_frames, n_frames, NULL, NULL))	/* same 2048-byte layout */
		return 1;
	if (!blocks)
		for (int i = 0; i < *n_frames; i++)
			frames[i].block = -1;
	return 0;
}

int vdl2_drain_frames_packed(vdl2gpu_t * h, vdl2_frame_hdr_t * hdrs, int max_frames, int *n_frames, uint8_t * bytes, size_t max_bytes, size_t *n_bytes,
			     vdl2_avlc_t * recs)
{				/* the same frames, in completion order, packed; field records are not produced by the stand-in */
	static vdl2_frame_t fr[4096];
	static vdl2_block_t bl[4096];
	int nf = 0, nb = 0;
	(void)recs;
	*n_frames = 0;
	*n_bytes = 0;
	if (vdl2_drain_frames(h, fr, 4096, &nf, bl, 4096, &nb) || nf > max_frames)
		return 1;
	int order[4096];
	for (int i = 0; i < nf; i++)
		order[i] = i;
	for (int i = 1; i < nf; i++) {	/* insertion sort by (end of burst, channel, length): a handful of frames per call */
		int k = order[i], j = i - 1;
		const int64_t ek = bl[fr[k].block].end_dump;
		while (j >= 0) {
			const vdl2_frame_t *a = &fr[order[j]];
			const int64_t ea = bl[a->block].end_dump;
			if (ea < ek || (ea == ek && (a->chn < fr[k].chn || (a->chn == fr[k].chn && a->len <= fr[k].len))))
				break;
			order[j + 1] = order[j];
			j--;
		}
		order[j + 1] = k;
	}
	size_t off = 0;
	for (int i = 0; i < nf; i++) {
		const vdl2_frame_t *f = &fr[order[i]];
		if (off + (size_t) ((f->len + 15) & ~15) > max_bytes)
			return 1;
		hdrs[i].sync_dump = f->sync_dump;
		hdrs[i].chn = f->chn;
		hdrs[i].Fr = f->Fr;
		hdrs[i].ppm = f->ppm;
		hdrs[i].len = f->len;
		hdrs[i].offset = (uint32_t) off;
		hdrs[i].dur = (int32_t) (bl[f->block].end_dump - f->sync_dump);
		memcpy(bytes + off, f->hdata, (size_t) f->len);
		off += (size_t) ((f->len + 15) & ~15);
	}
	*n_frames = nf;
	*n_bytes = off;
	return 0;
}

int vdl2_host_alloc(size_t bytes, void **out) { *out = malloc(bytes); return *out == NULL; }
int vdl2_host_free(void *p) { free(p); return 0; }
