/*
 * vdl2_host.cu -- host side of libvdl2gpu.so: the C ABI of include/vdl2gpu.h.
 *
 * Owns the device buffers (per-channel state, NCO tables, block queue, tap logs, the input
 * staging ring), builds the TMA descriptor for every launch and launches the fused
 * front-end kernel; behind it the block pipeline (vdl2_link.cu, row f1), the frame-field kernel (vdl2_avlc.cu,
 * row f4), page-locked ingest buffers and the on-device rtl.c block expansion (row f2).  No demodulation
 * happens on the host; if CUDA is unavailable every entry point fails (there is deliberately no CPU fallback).
 *
 * Host-side arithmetic that must equal the reference's:
 *   NCO table   d8psk.c:353-357   wf[n] = cexpf(-n * Fo' * I), Fo' rounded to float,
 *               evaluated here with the same glibc cexpf so the table is bit-identical;
 *   dump clock  d8psk.c:374-381   clk += 21; dump when clk >= SDRCLK  -> schedule table.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include "vdl2_kernel.h"
#include "vdl2_link.h"
#include "vdl2_tables.h"
#include "vdl2_mma_tables.h"

static_assert(sizeof(Vdl2BlockRec) == sizeof(vdl2_block_t), "block record layout");
static_assert(sizeof(Vdl2StepRec) == sizeof(vdl2_step_t), "step record layout");
static_assert(sizeof(Vdl2SyncRec) == sizeof(vdl2_sync_t), "sync record layout");
static_assert(sizeof(Vdl2SymRec) == sizeof(vdl2_sym_t), "sym record layout");
static_assert(sizeof(Vdl2FrameRec) == sizeof(vdl2_frame_t) && sizeof(vdl2_frame_t) == 2048, "frame record layout");
static_assert(sizeof(Vdl2BlkStat) == sizeof(vdl2_blkstat_t), "block statistics layout");

#define VDL2_PIN_SLOTS 4
#define VDL2_MIRROR_SLOTS 8	/* > VDL2_PIN_SLOTS: a vdl2_submit_copy() caller is never more launches ahead of the device than the ring is deep */
static thread_local std::string g_create_error;

struct vdl2gpu {
	vdl2_config_t cfg;
	int nstreams;
	int bytes_per_sample;	/* per IQ sample (or per real sample) */
	int row_samples, row_bytes, chunks_per_row, nbox, spc;
	int nco_entries, wext;
	int smem, grid, n_sm, ctas_per_sm;
	cudaStream_t stream;
	bool l2_manage;		/* this handle switches the device's persisting-L2 limit per path (l2_reserve) */
	bool l2_counted;	/* ... and is counted in g_l2_users */
	size_t l2_want;		/* set-aside of the fused kernel: scratch + mixer tables + 2 MB, capped by the device */
	cudaEvent_t ev0, ev1;
	bool ev_valid;
	/* device */
	Vdl2ChanState *d_state;
	float4 *d_wtab;
	float4 *d_dcorr;
	uint4 *d_w8;		/* integer mixer tables (dp4a mode only) */
	float *d_soft;		/* soft-demap tables for per-lane look-ups */
	int dp4a;		/* mixer of cu8/cs8 input at a rate whose dumps are 23/24 samples (2 Msps): 2 = int8 tensor cores (default),
				   1 = IDP.4A (VDL2_OPT_DP4A_MIX); 0 = the generic fp32 mixer (every other format / rate) */
	int sched_slot;
	unsigned *d_ticket;
	unsigned *d_slotmask;
	unsigned nsmid;
	int *d_progress;
	unsigned tiles_done;	/* per channel, since create (wraps; the kernel compares differences) */
	bool last_was_launch;	/* the newest operation on the stream is a front-end launch (programmatic dependent launch is safe behind it) */
	unsigned launch_seq;
	bool overlap;		/* VDL2_OPT_OVERLAP: no per-launch events, consecutive launches may overlap */
	uint8_t *d_curblk;
	float2 *d_scratch;
	Vdl2BlockRec *d_outq;
	unsigned *d_outq_count, *d_dropped;
	unsigned outq_cap;
	float2 *d_tap_dumps;
	Vdl2StepRec *d_tap_steps;
	Vdl2SyncRec *d_tap_syncs;
	Vdl2SymRec *d_tap_syms;
	unsigned cap_dumps, cap_steps, cap_syncs, cap_syms;
	uint8_t *d_stage;	/* [nstreams][stage_pitch]: carry + new samples */
	size_t stage_pitch;
	size_t carry;		/* samples pending at the head of each staging stream */
	int64_t rows_done;
	vdl2_stats_t st;
	std::string err;
	void *encode_fn;
	/* block pipeline (vdl2_link.cu), allocated on first use */
	Vdl2BlockRec *d_lblocks;
	Vdl2FrameRec *d_frames;
	Vdl2BlkStat *d_lstats;
	uint8_t *d_lrows;
	unsigned *d_nframes;
	int lcap_blocks, lcap_frames;
	cudaEvent_t lev0, lev1;
	bool lev_valid, link_ready;
	void *d_avlc;		/* field records of one vdl2_avlc_extract() call, grown on demand */
	int avlc_cap;
	/* packed drain (vdl2_drain_frames_packed): rank / offsets / headers / bytes / records on the device, totals in page-locked memory */
	int *d_rank;
	unsigned *d_offs, *d_totals;
	void *d_hdrs, *d_precs, *d_keys;
	uint8_t *d_pbytes;
	int pack_cap;
	unsigned *h_totals;
	cudaEvent_t pev0, pev1;
	float last_pack_ms;
	uint8_t *d_raw;		/* raw cu8 bytes of one vdl2_process_host_rtl() call, grown on demand */
	size_t raw_cap;
	/* asynchronous ingest (vdl2_submit_copy): ring of page-locked slots the caller's buffer is copied into, one event per slot
	   (recorded behind the slot's upload); mirror of the device counters in page-locked host memory, refreshed behind every launch */
	uint8_t *pin[VDL2_PIN_SLOTS];
	cudaEvent_t pin_ev[VDL2_PIN_SLOTS];
	size_t pin_bytes;
	unsigned pin_next;
	unsigned *h_mirror;	/* [VDL2_MIRROR_SLOTS][16] copies of d_ticket, one slot per launch in turn */
	cudaEvent_t mirror_ev[VDL2_MIRROR_SLOTS];
	unsigned mirror_seq;	/* launches whose mirror copy was enqueued */
};

static int fail(vdl2gpu * h, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	fprintf(stderr, "vdl2gpu: %s\n", buf);	/* reference convention: message on stderr, non-zero return */
	if (h)
		h->err = buf;
	else
		g_create_error = buf;
	return 1;
}

#define CK(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

/* persisting-L2 limit of the device as this library last set it (-1: not yet); one entry per device ordinal */
static std::atomic < long long >g_l2_limit[64];
static std::atomic < int >g_l2_users[64];	/* live handles per device that manage the limit: the last one out puts it back to 0 */
static struct L2LimitInit { L2LimitInit() { for (auto & v:g_l2_limit) v.store(-1); } } g_l2_limit_init;

/* the fused kernel wants its scratch and tables reserved in L2, the channeliser wants all of L2: switch the device limit when the
   path changes (a handful of times in the life of a process; a relaxed load per launch otherwise) */
static int l2_reserve(vdl2gpu * h, size_t bytes)
{
	if (!h->l2_manage || h->cfg.device < 0 || h->cfg.device >= 64)
		return 0;
	if (g_l2_limit[h->cfg.device].load(std::memory_order_relaxed) == (long long)bytes)
		return 0;
	CK(h, cudaStreamSynchronize(h->stream));
	/* an optimisation only: where the limit cannot be set (e.g. a context that does not own its L2), carry on without it */
	if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) {
		(void)cudaGetLastError();
		h->l2_manage = false;
		return 0;
	}
	if (bytes == 0 && cudaCtxResetPersistingL2Cache() != cudaSuccess)
		(void)cudaGetLastError();
	g_l2_limit[h->cfg.device].store((long long)bytes, std::memory_order_relaxed);
	return 0;
}


extern "C" int vdl2_abi_version(void)
{
	return VDL2_ABI_VERSION;
}

extern "C" const char *vdl2_last_error(const vdl2gpu_t * h)
{
	return h ? h->err.c_str() : g_create_error.c_str();
}

static int fmt_bytes(int fmt)
{
	switch (fmt) {
	case VDL2_FMT_CU8: case VDL2_FMT_CS8: return 2;
	case VDL2_FMT_CS16: return 4;
	case VDL2_FMT_CF32: return 8;
	case VDL2_FMT_F32REAL: return 4;
	}
	return 0;
}

static void build_tables(Vdl2Tables & t, unsigned *sched_dump, const vdl2gpu * h)
{
	memset(&t, 0, sizeof t);
	float mf[VDL2_MFLTLEN], sw[VDL2_NBPH];
	static float soft[3][257];
	vdl2_make_mflt(mf);
	vdl2_make_sync(sw);
	vdl2_make_softmap(soft);
	memcpy(t.mflt, mf, sizeof mf);
	memcpy(t.sync, sw, sizeof sw);
	for (int b = 0; b < 3; b++)
		memcpy(t.soft[b], soft[b], sizeof soft[b]);
	/* descrambler sequence (d8psk.c:54-65), seed 0x4D4B (d8psk.c:299) */
	unsigned s = 0x4D4B;
	for (int i = 0; i < VDL2_SCR_WORDS * 32; i++) {
		unsigned b = (s ^ (s >> 14)) & 1u;
		s = (s << 1) | b;
		t.scr[i >> 5] |= b << (i & 31);
	}
	static const unsigned char hc[25] = { 0x06, 0x07, 0x09, 0x0a, 0x0b, 0x0c, 0x0e, 0x0f, 0x11, 0x13, 0x15, 0x16, 0x18,
		0x19, 0x1a, 0x1b, 0x1c, 0x1d, 0x1e, 0x1f, 0x10, 0x08, 0x04, 0x02, 0x01
	};			/* viterbi.c:29-35 */
	memcpy(t.hcol, hc, sizeof hc);
	/* dump schedule of one row (d8psk.c:374-381): dump k ends after sample e_k; it consists of np whole
	   chunks plus the chunk holding e_k (np counts from the chunk after the previous boundary chunk) */
	int clk = 0, k = 0, prev_c = -1;
	if (h->dp4a == 2) {
		vdl2_mma_build_sched(h->row_samples, (int)h->cfg.sdrclk, h->cfg.fs / VDL2_STEPRATE, VDL2_DUMPS_PER_ROW, sched_dump);
		return;
	}
	if (h->dp4a) {
		/* integer mixer: per dump (first sample << 16) | (table entry of its 12th sample pair << 8) | entry of its
		   first pair.  Entries 0..103 are indexed by NCO phase; a 23-sample dump closes with one of the
		   "first sample only" entries 104.. (assigned in row order, the same order the table builder uses) */
		int start = 0, nshort = 0;
		const int nco_n = h->cfg.fs / VDL2_STEPRATE;
		for (int n = 0; n < h->row_samples; n++) {
			clk += 21;
			if (clk >= (int)h->cfg.sdrclk) {
				clk %= (int)h->cfg.sdrclk;
				const int len = n + 1 - start, w0 = start % nco_n;
				const int wl = (len == 24) ? w0 + 22 : VDL2_W8_PHASES + nshort++;
				if (k < VDL2_DUMPS_PER_ROW)
					sched_dump[k] = ((unsigned)start << 16) | ((unsigned)wl << 8) | (unsigned)w0;
				start = n + 1;
				k++;
			}
		}
		return;
	}
	for (int n = 0; n < h->row_samples; n++) {
		clk += 21;
		if (clk >= (int)h->cfg.sdrclk) {
			clk %= (int)h->cfg.sdrclk;
			const int c = n / h->spc, E = n % h->spc;
			const int wpc = (h->spc == 8) ? 4 : 2;
			const int w0 = ((prev_c + 1) * wpc) % h->nco_entries;
			if (k < VDL2_DUMPS_PER_ROW)
				sched_dump[k] = ((unsigned)w0 << 16) | ((unsigned)E << 8) | (unsigned)(c - prev_c - 1);
			prev_c = c;
			k++;
		}
	}
}

static int create_body(vdl2gpu * h, const vdl2_config_t * cfg, const vdl2_chan_param_t * chans, const cudaDeviceProp & prop);

extern "C" int vdl2_create(const vdl2_config_t * cfg, const vdl2_chan_param_t * chans, vdl2gpu_t ** out)
{
	if (!cfg || !chans || !out)
		return fail(NULL, "vdl2_create: null argument");
	*out = NULL;
	if (cfg->nch <= 0 || cfg->ch_per_stream <= 0 || cfg->nch % cfg->ch_per_stream)
		return fail(NULL, "vdl2_create: nch=%d must be a positive multiple of ch_per_stream=%d", cfg->nch, cfg->ch_per_stream);
	if (cfg->format < VDL2_FMT_CU8 || cfg->format > VDL2_FMT_F32REAL)
		return fail(NULL, "vdl2_create: unknown sample format %d", cfg->format);
	if (cfg->fs % 1000 || cfg->fs % VDL2_STEPRATE || cfg->sdrclk == 0 || (cfg->fs / 1000 * 21) % cfg->sdrclk)
		return fail(NULL, "vdl2_create: fs=%u / sdrclk=%u do not give a 1 ms joint period", cfg->fs, cfg->sdrclk);
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail(NULL, "vdl2_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
	if (cfg->device < 0 || cfg->device >= ndev)
		return fail(NULL, "vdl2_create: device %d out of range (%d present)", cfg->device, ndev);
	CK(NULL, cudaSetDevice(cfg->device));
	cudaDeviceProp prop;
	CK(NULL, cudaGetDeviceProperties(&prop, cfg->device));
	if (prop.major < 10)
		return fail(NULL, "vdl2_create: device %d is sm_%d%d; the kernels are built for sm_100a only", cfg->device, prop.major,
			    prop.minor);

	/* everything after this point can fail half way (out of memory, driver entry points ...): the partially built handle is
	   torn down by the same code as a complete one, and its message moves to the create-error slot the caller can read
	   through vdl2_last_error(NULL) */
	vdl2gpu *h = new vdl2gpu();	/* value-initialised: every pointer null, every flag false */
	const int rc = create_body(h, cfg, chans, prop);
	if (rc) {
		g_create_error = h->err;
		vdl2_destroy(h);
		return rc;
	}
	*out = h;
	return 0;
}

static int create_body(vdl2gpu * h, const vdl2_config_t * cfg, const vdl2_chan_param_t * chans, const cudaDeviceProp & prop)
{
	cudaError_t e = cudaSuccess;
	h->cfg = *cfg;
	h->nstreams = cfg->nch / cfg->ch_per_stream;
	h->bytes_per_sample = fmt_bytes(cfg->format);
	h->row_samples = cfg->fs / 1000;
	h->row_bytes = h->row_samples * h->bytes_per_sample;
	h->spc = 16 / h->bytes_per_sample;
	h->ev_valid = false;
	h->d_raw = NULL;
	h->raw_cap = 0;
	h->d_avlc = NULL;
	h->avlc_cap = 0;
	h->d_lblocks = NULL;
	h->d_frames = NULL;
	h->d_lstats = NULL;
	h->d_lrows = NULL;
	h->d_nframes = NULL;
	h->lcap_blocks = h->lcap_frames = 0;
	h->lev_valid = h->link_ready = false;
	h->carry = 0;
	h->rows_done = 0;
	memset(&h->st, 0, sizeof h->st);
	if (h->row_bytes % 16 || h->row_bytes / 16 > VDL2_MAX_CHUNKS) {
		return fail(h, "vdl2_create: row of %d bytes is not a whole number of 16-byte chunks", h->row_bytes);
	}
	h->chunks_per_row = h->row_bytes / 16;
	h->nbox = (h->row_bytes + 127) / 128;	/* generic mixer: 128-byte column boxes; the tensor-core mixer sets its own below */
	const int nco_n = cfg->fs / VDL2_STEPRATE;	/* d8psk.c:348 */
	h->nco_entries = (cfg->format == VDL2_FMT_CF32) ? nco_n : nco_n / 2;
	h->n_sm = prop.multiProcessorCount;
	/* integer mixer: 8-bit input whose dumps are all 23 or 24 samples long (2 Msps, SDRCLK 500: rtl.c:36-37) */
	h->dp4a = 0;
	h->d_w8 = NULL;
	if ((cfg->format == VDL2_FMT_CU8 || cfg->format == VDL2_FMT_CS8) && !(cfg->taps & VDL2_OPT_FLOAT_MIX) && nco_n + 24 <= VDL2_W8_PHASES
	    && !getenv("VDL2_FLOAT_MIX")) {
		int clk = 0, start = 0, ok = 1, nshort = 0;
		for (int n = 0; n < h->row_samples; n++) {
			clk += 21;
			if (clk >= (int)cfg->sdrclk) {
				clk %= (int)cfg->sdrclk;
				const int len = n + 1 - start;
				if (len != 23 && len != 24)
					ok = 0;
				nshort += (len == 23);
				start = n + 1;
			}
		}
		h->dp4a = ok && nshort <= VDL2_W8_ENTRIES - VDL2_W8_PHASES;
		/* the tensor-core mixer unless the caller asks for the IDP.4A one (A/B) */
		if (h->dp4a && !(cfg->taps & VDL2_OPT_DP4A_MIX) && !getenv("VDL2_DP4A_MIX")
		    && vdl2_mma_usable(h->row_samples, (int)cfg->sdrclk, nco_n, VDL2_DUMPS_PER_ROW, VDL2_MM_PHASES))
			h->dp4a = 2;
	}

	/* dump schedule sanity: exactly 84 dumps per row, clock back at 0, <= 1 boundary per chunk */
	{
		int clk = 0, k = 0, lastc = -1;
		for (int n = 0; n < h->row_samples; n++) {
			clk += 21;
			if (clk >= (int)cfg->sdrclk) {
				clk %= (int)cfg->sdrclk;
				k++;
				if (n / h->spc == lastc) {
					return fail(h, "vdl2_create: two dump boundaries in one chunk (fs too low)");
				}
				lastc = n / h->spc;
			}
		}
		if (k != VDL2_DUMPS_PER_ROW || clk != 0) {
			return fail(h, "vdl2_create: fs=%u sdrclk=%u gives %d dumps/ms (clk %d), need 84 (0)", cfg->fs, cfg->sdrclk, k, clk);
		}
	}

	Vdl2Tables *tab = new Vdl2Tables();
	unsigned sched_dump[VDL2_DUMPS_PER_ROW];
	build_tables(*tab, sched_dump, h);
	e = (cudaError_t) vdl2_kernel_upload_tables(tab);
	h->d_soft = NULL;
	if (e == cudaSuccess)
		e = cudaMalloc(&h->d_soft, sizeof tab->soft + sizeof tab->scr);
	if (e == cudaSuccess)
		e = cudaMemcpy(h->d_soft, tab->soft, sizeof tab->soft, cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy((char *)h->d_soft + sizeof tab->soft, tab->scr, sizeof tab->scr, cudaMemcpyHostToDevice);
	delete tab;
	if (e == cudaSuccess) {
		/* the dump schedule depends on (fs, SDRCLK, format): handles with the same signature share a
		   constant-memory slot (identical content), up to VDL2_SCHED_SLOTS signatures per device */
		static std::mutex mtx;
		static struct { unsigned fs, sdrclk; int fmt, dev, dp; } slots[VDL2_SCHED_SLOTS];
		static int nslots = 0;
		std::lock_guard < std::mutex > lk(mtx);
		h->sched_slot = -1;
		for (int i = 0; i < nslots; i++)
			if (slots[i].fs == cfg->fs && slots[i].sdrclk == cfg->sdrclk && slots[i].fmt == cfg->format && slots[i].dev == cfg->device
			    && slots[i].dp == h->dp4a)
				h->sched_slot = i;
		if (h->sched_slot < 0) {
			if (nslots == VDL2_SCHED_SLOTS) {
				return fail(h, "vdl2_create: more than %d distinct (fs, sdrclk, format) combinations in one process", VDL2_SCHED_SLOTS);
			}
			slots[nslots].fs = cfg->fs;
			slots[nslots].sdrclk = cfg->sdrclk;
			slots[nslots].fmt = cfg->format;
			slots[nslots].dev = cfg->device;
			slots[nslots].dp = h->dp4a;
			h->sched_slot = nslots++;
		}
		e = (cudaError_t) vdl2_kernel_upload_sched(h->sched_slot, sched_dump);
	}
	if (e != cudaSuccess) {
		return fail(h, "vdl2_create: constant table upload failed: %s", cudaGetErrorString(e));
	}

	/* longest dump in chunks (+1 boundary chunk) times entries per chunk: how far past the table a dump can read */
	h->wext = ((int)((cfg->fs / 84000 + 2 + h->spc - 1) / h->spc) + 1) * ((h->spc == 8) ? 4 : 2);
	h->smem = vdl2_kernel_smem_bytes(h->nco_entries + h->wext, h->dp4a);
	e = (cudaError_t) vdl2_kernel_occupancy(cfg->format, h->dp4a, h->smem, &h->ctas_per_sm);
	if (e != cudaSuccess || h->ctas_per_sm < 1) {
		return fail(h, "vdl2_create: kernel does not fit (smem %d B): %s", h->smem, cudaGetErrorString(e));
	}
	h->grid = h->n_sm * h->ctas_per_sm;
	h->st.n_sm = h->n_sm;
	h->st.smem_bytes = h->smem;

	CK(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	CK(h, cudaEventCreate(&h->ev0));
	CK(h, cudaEventCreate(&h->ev1));

	const int nch = cfg->nch;
	/* per-channel state: zero, then what initD8psk / initVdlm2 set (d8psk.c:28-37, vdlm2.c:167) */
	std::vector < Vdl2ChanState > st(nch);
	memset(st.data(), 0, sizeof(Vdl2ChanState) * nch);
	for (int c = 0; c < nch; c++) {
		st[c].perr = 100.f;
		st[c].state = VDL2_ST_WSYNC;
		st[c].chn = chans[c].chn;
		st[c].Fr = chans[c].Fr;
		st[c].sync_dump = -1;
	}
	CK(h, cudaMalloc(&h->d_state, sizeof(Vdl2ChanState) * nch));
	CK(h, cudaMemcpy(h->d_state, st.data(), sizeof(Vdl2ChanState) * nch, cudaMemcpyHostToDevice));

	/* NCO tables, d8psk.c:353-357 */
	std::vector < float4 > wt((size_t) nch * h->nco_entries);
	for (int c = 0; c < nch; c++) {
		/* cexpf(-n*Fo*I) = (cosf(a), sinf(a)) with the float product a = -n*Fo: vdl2_nco_table (vdl2_mma_tables.h) is the one
		   implementation every mixer table is built from; tests/test_oracle.py::test_nco_table_bit_identical_to_reference pins
		   it against the reference's cexpf table over the whole 25 kHz raster at every supported rate */
		std::vector < float >wr(nco_n), wi(nco_n);
		vdl2_nco_table(chans[c].Fo, cfg->fs, nco_n, wr.data(), wi.data());
		float4 *dst = wt.data() + (size_t) c * h->nco_entries;
		if (cfg->format == VDL2_FMT_CF32) {
			for (int n = 0; n < nco_n; n++)
				dst[n] = make_float4(wr[n], wr[n], wi[n], wi[n]);
		} else {
			for (int p = 0; p < nco_n / 2; p++)
				dst[p] = make_float4(wr[2 * p], wr[2 * p + 1], wi[2 * p], wi[2 * p + 1]);
		}
	}
	CK(h, cudaMalloc(&h->d_wtab, sizeof(float4) * wt.size()));
	CK(h, cudaMemcpy(h->d_wtab, wt.data(), sizeof(float4) * wt.size(), cudaMemcpyHostToDevice));

	/* per channel, per dump of a row: 1/nf and the cu8 offset correction.  The kernel mixes the exact
	   integers u-127; the reference mixes u-127.37f (rtl.c:287-289), i.e. (u-127) - delta with
	   delta = 127.37f - 127 (exact in fp32).  sum (x-delta) w = sum x w - delta sum w, so per dump
	   re -= delta (sum wr - sum wi), im -= delta (sum wr + sum wi), evaluated here in double. */
	{
		std::vector < float4 > dc((size_t) nch * VDL2_DUMPS_PER_ROW);
		const double delta = (cfg->format == VDL2_FMT_CU8) ? (double)((float)127.37 - 127.0f) : 0.0;
		for (int c = 0; c < nch; c++) {
			const float Fo = (float)((float)chans[c].Fo / (float)(cfg->fs) * 2.0 * M_PI);
			int clk = 0, nf = 0, k = 0;
			double swr = 0, swi = 0;
			for (int n = 0; n < h->row_samples; n++) {
				const float a = (float)(-(n % nco_n)) * Fo;
				swr += (double)cosf(a);
				swi += (double)sinf(a);
				nf++;
				clk += 21;
				if (clk >= (int)cfg->sdrclk) {
					clk %= (int)cfg->sdrclk;
					const double s = 1.0 / (double)nf;
					const float sf = 1.0f / (float)nf;
					if (k < VDL2_DUMPS_PER_ROW)
						dc[(size_t) c * VDL2_DUMPS_PER_ROW + k] =
						    make_float4(sf, sf, (float)(-delta * (swr - swi) * s), (float)(-delta * (swr + swi) * s));
					k++;
					nf = 0;
					swr = swi = 0;
				}
			}
		}
		CK(h, cudaMalloc(&h->d_dcorr, sizeof(float4) * dc.size()));
		CK(h, cudaMemcpy(h->d_dcorr, dc.data(), sizeof(float4) * dc.size(), cudaMemcpyHostToDevice));
	}

	if (h->dp4a == 2) {
		/* tensor-core mixer: B fragments per window phase and per-dump accumulator start / scale / correction
		   (vdl2_mma_tables.h); they travel in the w8 / dcorr slots of the kernel parameters */
		static_assert(sizeof(Vdl2MmaU4) == sizeof(uint4) && sizeof(Vdl2MmaI4) == sizeof(float4), "table entry layout");
		std::vector < Vdl2MmaU4 > bt((size_t) nch * VDL2_MM_BT_ENTRIES);
		std::vector < Vdl2MmaI4 > dt((size_t) nch * VDL2_MM_DT_ENTRIES);
		std::vector < float >wr(nco_n), wi(nco_n);
		for (int c = 0; c < nch; c++) {
			vdl2_nco_table(chans[c].Fo, cfg->fs, nco_n, wr.data(), wi.data());
			vdl2_mma_build_chan(wr.data(), wi.data(), nco_n, h->row_samples, (int)cfg->sdrclk, VDL2_DUMPS_PER_ROW, cfg->format == VDL2_FMT_CU8,
					    bt.data() + (size_t) c * VDL2_MM_BT_ENTRIES, dt.data() + (size_t) c * VDL2_MM_DT_ENTRIES);
		}
		unsigned sched_tmp[VDL2_DUMPS_PER_ROW];
		h->nbox = vdl2_mma_build_sched(h->row_samples, (int)cfg->sdrclk, nco_n, VDL2_DUMPS_PER_ROW, sched_tmp);
		CK(h, cudaMalloc(&h->d_w8, sizeof(uint4) * bt.size()));
		CK(h, cudaMemcpy(h->d_w8, bt.data(), sizeof(uint4) * bt.size(), cudaMemcpyHostToDevice));
		CK(h, cudaFree(h->d_dcorr));
		h->d_dcorr = NULL;
		CK(h, cudaMalloc(&h->d_dcorr, sizeof(float4) * dt.size()));
		CK(h, cudaMemcpy(h->d_dcorr, dt.data(), sizeof(float4) * dt.size(), cudaMemcpyHostToDevice));
	} else if (h->dp4a) {
		/* integer mixer tables.  W = round(w * 2^22) of the reference's float oscillator value, split into balanced
		   base-256 digits W = d2*65536 + d1*256 + d0 (each in [-128,127], |d2| <= 64).  The kernel accumulates
		     re_acc = sum Is*wr + (~Qs)*wi = sum Is*wr - Qs*wi - sum wi,    im_acc = sum Qs*wr + Is*wi
		   with Is = I - 128 (cu8) or I (cs8), while the reference mixes x = Is + d, d = 128 - 127.37f (cu8) or 0, so
		     re = re_acc + sum wi + d (sum wr - sum wi),   im = im_acc + d (sum wr + sum wi)
		   (the "+ sum wi" uses the QUANTISED weights: it undoes an exact integer identity). */
		std::vector < uint4 > w8((size_t) nch * VDL2_W8_ENTRIES);
		std::vector < float4 > dc((size_t) nch * VDL2_DUMPS_PER_ROW);
		const double d = (cfg->format == VDL2_FMT_CU8) ? 128.0 - (double)(float)127.37 : 0.0;
		for (int c = 0; c < nch; c++) {
			const float Fo = (float)((float)chans[c].Fo / (float)(cfg->fs) * 2.0 * M_PI);
			std::vector < int >qr(nco_n), qi(nco_n);
			std::vector < double >fr(nco_n), fi(nco_n);
			for (int n = 0; n < nco_n; n++) {
				const float a = (float)(-n) * Fo;
				fr[n] = (double)cosf(a);
				fi[n] = (double)sinf(a);
				qr[n] = (int)lrint(fr[n] * 4194304.0);
				qi[n] = (int)lrint(fi[n] * 4194304.0);
			}
			auto digits =[](int W, int dg[3]) {
				dg[0] = ((W + 128) & 255) - 128;
				W = (W - dg[0]) / 256;
				dg[1] = ((W + 128) & 255) - 128;
				dg[2] = (W - dg[1]) / 256;
			};
			auto entry =[&](int n, bool first_only) {
				int a[3], b[3], a1[3] = { 0, 0, 0 }, b1[3] = { 0, 0, 0 };
				digits(qr[n % nco_n], a);
				digits(qi[n % nco_n], b);
				if (!first_only) {
					digits(qr[(n + 1) % nco_n], a1);
					digits(qi[(n + 1) % nco_n], b1);
				}
				unsigned w[3];
				for (int j = 0; j < 3; j++)
					w[j] = (unsigned)(a[j] & 255) | ((unsigned)(b[j] & 255) << 8) | ((unsigned)(a1[j] & 255) << 16) | ((unsigned)(b1[j] & 255) << 24);
				return make_uint4(w[2], w[1], w[0], 0u);
			};
			uint4 *wt = w8.data() + (size_t) c * VDL2_W8_ENTRIES;
			for (int n = 0; n < VDL2_W8_PHASES; n++)
				wt[n] = entry(n, false);
			int clk = 0, start = 0, k = 0, nshort = 0;
			long long sqi = 0;
			double swr = 0, swi = 0;
			for (int n = 0; n < h->row_samples; n++) {
				sqi += qi[n % nco_n];
				swr += fr[n % nco_n];
				swi += fi[n % nco_n];
				clk += 21;
				if (clk >= (int)cfg->sdrclk) {
					clk %= (int)cfg->sdrclk;
					const int len = n + 1 - start;
					if (len == 23)
						wt[VDL2_W8_PHASES + nshort++] = entry(n, true);	/* the dump's last sample alone */
					const double s = 1.0 / (double)len, q = 1.0 / 4194304.0;
					const float sf = (float)(q * s);
					if (k < VDL2_DUMPS_PER_ROW)
						dc[(size_t) c * VDL2_DUMPS_PER_ROW + k] =
						    make_float4(sf, sf, (float)(((double)sqi * q + d * (swr - swi)) * s), (float)((d * (swr + swi)) * s));
					k++;
					start = n + 1;
					sqi = 0;
					swr = swi = 0;
				}
			}
		}
		CK(h, cudaMalloc(&h->d_w8, sizeof(uint4) * w8.size()));
		CK(h, cudaMemcpy(h->d_w8, w8.data(), sizeof(uint4) * w8.size(), cudaMemcpyHostToDevice));
		CK(h, cudaMemcpy(h->d_dcorr, dc.data(), sizeof(float4) * dc.size(), cudaMemcpyHostToDevice));
	}

	CK(h, cudaMalloc(&h->d_ticket, 256));	/* words 16..: chain statistics of a -DVDL2_CHAIN_STATS build */
	CK(h, cudaMemset(h->d_ticket, 0, 256));
	h->d_outq_count = h->d_ticket + 4;
	h->d_dropped = h->d_ticket + 8;
	CK(h, cudaMalloc(&h->d_progress, sizeof(int) * nch));
	CK(h, cudaMemset(h->d_progress, 0, sizeof(int) * nch));
	h->tiles_done = 0;
	h->launch_seq = 0;
	h->overlap = (cfg->taps & VDL2_OPT_OVERLAP) != 0;
	h->nsmid = 0;
	if (vdl2_kernel_nsmid(h->d_ticket + 15, &h->nsmid) || h->nsmid == 0 || h->nsmid > 1024)
		return fail(h, "vdl2_create: cannot query the SM id range");
	CK(h, cudaMemset(h->d_ticket + 15, 0, 4));
	CK(h, cudaMalloc(&h->d_slotmask, sizeof(unsigned) * h->nsmid));
	CK(h, cudaMemset(h->d_slotmask, 0, sizeof(unsigned) * h->nsmid));
	CK(h, cudaMalloc(&h->d_curblk, (size_t) nch * 2048));
	CK(h, cudaMemset(h->d_curblk, 0, (size_t) nch * 2048));
	{	/* one scratch slot per (SM id, resident CTA); SM ids are not contiguous on parts with disabled SMs */
		const size_t nslots = (size_t) h->nsmid * h->ctas_per_sm;
		CK(h, cudaMalloc(&h->d_scratch, sizeof(float2) * nslots * (VDL2_HIST + VDL2_TILE_DUMPS)));
		CK(h, cudaMemset(h->d_scratch, 0, sizeof(float2) * nslots * (VDL2_HIST + VDL2_TILE_DUMPS)));
		/* L2 set-aside for what the fused kernel keeps on chip: the per-warp scratch (written in phase 1, read back in phase 2) and
		   the per-channel mixer tables carry evict_last hints, but without a reserved share of L2 the 8.6 GB input stream still pushed
		   0.3-0.45 GB of scratch per step out to DRAM (ncu, profiles/r2_ab_l2persist.txt: write-backs 0.3 -> 0.02 GB with the
		   set-aside, noise probe 2.5 % faster).  The limit belongs to the device context, not to this handle, and the one-pass
		   channeliser (row f3), which keeps nothing on chip, runs 2.2 x SLOWER with half of L2 reserved: so the limit is switched
		   per path (l2_reserve below: a device call only when the path changes).  VDL2_L2_PERSIST_MB=0 never touches the
		   limit, =<n> asks for n MB. */
		const size_t bytes = sizeof(float2) * nslots * (VDL2_HIST + VDL2_TILE_DUMPS);
		const char *pe = getenv("VDL2_L2_PERSIST_MB");
		size_t want = bytes + (h->dp4a == 2 ? (size_t) nch * (VDL2_MM_BT_ENTRIES * 16 + VDL2_MM_DT_ENTRIES * 16) : 0) + ((size_t) 2 << 20);
		if (pe)
			want = (size_t) atoi(pe) << 20;
		h->l2_manage = !(pe && atoi(pe) == 0);
		if (h->l2_manage && cfg->device >= 0 && cfg->device < 64) {
			g_l2_users[cfg->device].fetch_add(1);
			h->l2_counted = true;
		}
		h->l2_want = std::min(want, (size_t) prop.persistingL2CacheMaxSize);
		if (getenv("VDL2_PRE_STATS"))
			fprintf(stderr, "vdl2gpu: L2 set-aside %zu MB (device maximum %d MB), scratch %zu MB\n", h->l2_manage ? h->l2_want >> 20 : (size_t) 0,
				prop.persistingL2CacheMaxSize >> 20, bytes >> 20);
	}
	h->outq_cap = cfg->max_blocks > 0 ? (unsigned)cfg->max_blocks : (unsigned)std::max(4096, nch * 8);
	CK(h, cudaMalloc(&h->d_outq, sizeof(Vdl2BlockRec) * (size_t) h->outq_cap));

	const size_t maxs = cfg->max_samples ? cfg->max_samples : (size_t) 1 << 20;
	const size_t max_rows = maxs / h->row_samples + 2;
	h->cap_dumps = h->cap_steps = h->cap_syncs = h->cap_syms = 0;
	h->d_tap_dumps = NULL;
	h->d_tap_steps = NULL;
	h->d_tap_syncs = NULL;
	h->d_tap_syms = NULL;
	if (cfg->taps & VDL2_TAP_DUMPS) {
		h->cap_dumps = (unsigned)(max_rows * VDL2_DUMPS_PER_ROW);
		CK(h, cudaMalloc(&h->d_tap_dumps, sizeof(float2) * (size_t) h->cap_dumps * nch));
	}
	if (cfg->taps & VDL2_TAP_STEPS) {
		h->cap_steps = (unsigned)(max_rows * VDL2_DUMPS_PER_ROW / 2 + 64);
		CK(h, cudaMalloc(&h->d_tap_steps, sizeof(Vdl2StepRec) * (size_t) h->cap_steps * nch));
	}
	if (cfg->taps & VDL2_TAP_SYNCS) {
		h->cap_syncs = (unsigned)(max_rows * VDL2_DUMPS_PER_ROW / 64 + 64);
		CK(h, cudaMalloc(&h->d_tap_syncs, sizeof(Vdl2SyncRec) * (size_t) h->cap_syncs * nch));
	}
	if (cfg->taps & VDL2_TAP_SYMS) {
		h->cap_syms = (unsigned)(max_rows * VDL2_DUMPS_PER_ROW / 8 + 64);
		CK(h, cudaMalloc(&h->d_tap_syms, sizeof(Vdl2SymRec) * (size_t) h->cap_syms * nch));
	}

	/* staging: room for a carried tail (< 1 row) plus one call */
	h->stage_pitch = ((maxs + h->row_samples) * h->bytes_per_sample + 255) & ~(size_t) 255;
	CK(h, cudaMalloc(&h->d_stage, h->stage_pitch * h->nstreams));

	cudaDriverEntryPointQueryResult qres;
	h->encode_fn = NULL;
	CK(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->encode_fn, cudaEnableDefault, &qres));
	if (!h->encode_fn || qres != cudaDriverEntryPointSuccess)
		return fail(h, "vdl2_create: cuTensorMapEncodeTiled not available from the driver");
	return 0;
}

extern "C" int vdl2_destroy(vdl2gpu_t * h)
{
	if (!h)
		return 0;
	cudaSetDevice(h->cfg.device);
	if (h->stream)
		cudaStreamSynchronize(h->stream);
	if (h->l2_counted && h->cfg.device >= 0 && h->cfg.device < 64 && g_l2_users[h->cfg.device].fetch_sub(1) == 1 && h->stream)
		(void)l2_reserve(h, 0);	/* the device limit goes back to what a fresh context has */
	cudaFree(h->d_state);
	cudaFree(h->d_wtab);
	cudaFree(h->d_dcorr);
	cudaFree(h->d_w8);
	cudaFree(h->d_soft);
	cudaFree(h->d_ticket);
	cudaFree(h->d_progress);
	cudaFree(h->d_slotmask);
	cudaFree(h->d_curblk);
	cudaFree(h->d_scratch);
	cudaFree(h->d_outq);
	cudaFree(h->d_tap_dumps);
	cudaFree(h->d_tap_steps);
	cudaFree(h->d_tap_syncs);
	cudaFree(h->d_tap_syms);
	cudaFree(h->d_stage);
	cudaFree(h->d_raw);
	cudaFree(h->d_avlc);
	cudaFree(h->d_lblocks);
	cudaFree(h->d_frames);
	cudaFree(h->d_lstats);
	cudaFree(h->d_lrows);
	cudaFree(h->d_nframes);
	cudaFree(h->d_rank);
	cudaFree(h->d_offs);
	cudaFree(h->d_totals);
	cudaFree(h->d_hdrs);
	cudaFree(h->d_precs);
	cudaFree(h->d_keys);
	cudaFree(h->d_pbytes);
	if (h->h_totals)
		cudaFreeHost(h->h_totals);
	if (h->pev0)
		cudaEventDestroy(h->pev0);
	if (h->pev1)
		cudaEventDestroy(h->pev1);
	for (int i = 0; i < VDL2_PIN_SLOTS; i++) {
		if (h->pin[i])
			cudaFreeHost(h->pin[i]);
		if (h->pin_ev[i])
			cudaEventDestroy(h->pin_ev[i]);
	}
	if (h->h_mirror)
		cudaFreeHost(h->h_mirror);
	for (int i = 0; i < VDL2_MIRROR_SLOTS; i++)
		if (h->mirror_ev[i])
			cudaEventDestroy(h->mirror_ev[i]);
	if (h->link_ready) {
		cudaEventDestroy(h->lev0);
		cudaEventDestroy(h->lev1);
	}
	if (h->ev0)
		cudaEventDestroy(h->ev0);
	if (h->ev1)
		cudaEventDestroy(h->ev1);
	if (h->stream)
		cudaStreamDestroy(h->stream);
	delete h;
	return 0;
}

typedef CUresult(*encode_tiled_t) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
				   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
				   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

/* demodulate `nrows` complete rows starting at `base` (device), streams `pitch` bytes apart */
static int run_rows(vdl2gpu * h, const void *base, size_t pitch, int nrows)
{
	if (nrows <= 0)
		return 0;
	if (((uintptr_t) base & 15) || (pitch & 15))
		return fail(h, "input base/pitch must be 16-byte aligned for TMA (base %p pitch %zu)", base, pitch);
	CUtensorMap tmap;
	const cuuint64_t strides[2] = { (cuuint64_t) h->row_bytes, (cuuint64_t) pitch };
	const cuuint32_t estr[3] = { 1, 1, 1 };
	CUresult r;
	if (h->dp4a) {
		/* one element = one IQ sample (2 bytes); a box = 32 samples x 32 rows from the 16-byte boundary at or before the
		   start of a dump (TMA boxes must start 16-byte aligned), 64B-swizzled; samples past the end of a row read as zero */
		const cuuint64_t dims[3] = { (cuuint64_t) h->row_samples, (cuuint64_t) nrows, (cuuint64_t) h->nstreams };
		const cuuint32_t box[3] = { 32, 32, 1 };
		/* L2 promotion 128 B: measured DRAM traffic per bench step 9.07 GB (1.056 x algorithmic) against 9.61 GB with 256 B, 9.22 GB
		   without and 9.05 GB with 64 B (the latter 6 % slower); profiles/r2_promo_ab.txt */
		CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
		if (const char *pe = getenv("VDL2_TMA_PROMO"))	/* A/B: 0 none, 1 64 B, 2 128 B, 3 256 B */
			promo = (CUtensorMapL2promotion) atoi(pe);
		r = ((encode_tiled_t) h->encode_fn) (&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void *)base, dims, strides, box, estr,
						    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	} else {
		const cuuint64_t dims[3] = { (cuuint64_t) (h->row_bytes / 4), (cuuint64_t) nrows, (cuuint64_t) h->nstreams };
		const cuuint32_t box[3] = { 32, 32, 1 };
		r = ((encode_tiled_t) h->encode_fn) (&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base, dims, strides, box, estr,
						    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
						    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	}
	if (r != CUDA_SUCCESS)
		return fail(h, "cuTensorMapEncodeTiled failed (%d): rows=%d row_bytes=%d pitch=%zu", (int)r, nrows, h->row_bytes, pitch);

	Vdl2KParams kp;
	memset(&kp, 0, sizeof kp);
	kp.nch = h->cfg.nch;
	kp.ch_per_stream = h->cfg.ch_per_stream;
	kp.nrows = nrows;
	kp.ntiles = (nrows + VDL2_ROWS_PER_TILE - 1) / VDL2_ROWS_PER_TILE;
	kp.chunks_per_row = h->chunks_per_row;
	kp.nbox = h->nbox;
	kp.nco_pairs = h->nco_entries;
	kp.wext = h->wext;
	kp.dump_base = h->rows_done * VDL2_DUMPS_PER_ROW;
	kp.state = h->d_state;
	kp.wtab = h->d_wtab;
	kp.dcorr = h->d_dcorr;
	kp.w8 = h->d_w8;
	kp.soft = h->d_soft;
	kp.scr = (const unsigned *)((const char *)h->d_soft + sizeof(((Vdl2Tables *) 0)->soft));
	kp.sched_slot = h->sched_slot;
	kp.ticket = h->d_ticket;
	kp.ticket_sel = (int)(h->launch_seq % 3u);
	kp.launch_seq = (int)h->launch_seq;
	kp.tile_base = (int)h->tiles_done;
	kp.progress = h->d_progress;
	kp.slotmask = h->d_slotmask;
	kp.slots_per_sm = h->ctas_per_sm;
	kp.curblk = h->d_curblk;
	kp.scratch = h->d_scratch;
	kp.outq = h->d_outq;
	kp.outq_count = h->d_outq_count;
	kp.outq_cap = h->outq_cap;
	kp.dropped = h->d_dropped;
	kp.taps = h->cfg.taps & 15u;
	kp.flags = (h->cfg.taps & VDL2_OPT_EXACT_IDLE) ? VDL2_FLAG_NO_SCREEN : 0u;
	if (getenv("VDL2_NO_PREPASS"))	/* A/B switch for tools/ab_probe.sh */
		kp.flags |= VDL2_FLAG_NO_PREPASS;
	kp.tap_dumps = h->d_tap_dumps;
	kp.tap_steps = h->d_tap_steps;
	kp.tap_syncs = h->d_tap_syncs;
	kp.tap_syms = h->d_tap_syms;
	kp.cap_dumps = h->cap_dumps;
	kp.cap_steps = h->cap_steps;
	kp.cap_syncs = h->cap_syncs;
	kp.cap_syms = h->cap_syms;


	if (getenv("VDL2_PRE_STATS")) {	/* debug: outcome counters of the speculative pass A of the previous launch */
		unsigned c[4];
		cudaStreamSynchronize(h->stream);
		cudaMemcpy(c, h->d_ticket + 12, sizeof c, cudaMemcpyDeviceToHost);
		fprintf(stderr, "vdl2gpu: prepass used %u wasted %u idle-without %u burst-start %u\n", c[0], c[1], c[2], c[3]);
		cudaMemset(h->d_ticket + 12, 0, sizeof c);
		unsigned cs[28];
		cudaMemcpy(cs, h->d_ticket + 16, sizeof cs, cudaMemcpyDeviceToHost);
		static const char *kinds[7] = { "idle, speculation used", "idle, pass A on the chain", "idle -> trigger -> in burst", "idle -> whole burst -> idle",
			"in burst -> in burst", "in burst -> idle", "in burst -> idle -> next trigger" };
		for (int k = 0; k < 7; k++)
			if (cs[4 * k]) {
				unsigned long long ns;
				memcpy(&ns, cs + 4 * k + 2, 8);
				fprintf(stderr, "vdl2gpu: chain %-32s %7u tiles  %8.2f us each  %10.1f us total\n", kinds[k], cs[4 * k], ns * 1e-3 / cs[4 * k], ns * 1e-3);
			}
		if (cs[0]) {
			unsigned long long part[3];
			cudaMemcpy(part, h->d_ticket + 48, sizeof part, cudaMemcpyDeviceToHost);
			fprintf(stderr, "vdl2gpu: chain idle step = load %.2f + demodulate %.2f + store/fence %.2f us\n", part[0] * 1e-3 / cs[0], part[1] * 1e-3 / cs[0],
				part[2] * 1e-3 / cs[0]);
			cudaMemset(h->d_ticket + 48, 0, sizeof part);
			unsigned long long sub[4];	/* of "demodulate": head of the tile again with the real history | exact fit of the candidates | candidates | tiles */
			cudaMemcpy(sub, h->d_ticket + 54, sizeof sub, cudaMemcpyDeviceToHost);
			if (sub[3])
				fprintf(stderr, "vdl2gpu: chain idle step, demodulate = head %.2f + exact fits %.2f us (%.1f candidates per tile)\n", sub[0] * 1e-3 / sub[3],
					sub[1] * 1e-3 / sub[3], (double)sub[2] / sub[3]);
			cudaMemset(h->d_ticket + 54, 0, sizeof sub);
		}
		cudaMemset(h->d_ticket + 16, 0, sizeof cs);
	}

	const long long items = (long long)kp.ntiles * kp.nch;
	const int grid = (int)std::min < long long >(items, h->grid);
	if (!h->overlap)	/* an event between two kernels would serialise them */
		CK(h, cudaEventRecord(h->ev0, h->stream));
	if (l2_reserve(h, h->l2_want))
		return -1;
	cudaError_t e = (cudaError_t) vdl2_kernel_launch(h->cfg.format, h->dp4a, &tmap, &kp, grid, h->smem, h->stream, h->overlap && h->last_was_launch && base != (const void *)h->d_stage);
	if (e != cudaSuccess)
		return fail(h, "kernel launch failed: %s", cudaGetErrorString(e));
	if (!h->overlap) {
		CK(h, cudaEventRecord(h->ev1, h->stream));
		h->ev_valid = true;
	}
	if (h->h_mirror) {	/* somebody polls (vdl2_pending_blocks): refresh the host mirror of the counters behind this launch */
		const unsigned q = h->mirror_seq % VDL2_MIRROR_SLOTS;
		CK(h, cudaMemcpyAsync(h->h_mirror + 16 * q, h->d_ticket, 64, cudaMemcpyDeviceToHost, h->stream));
		CK(h, cudaEventRecord(h->mirror_ev[q], h->stream));
		h->mirror_seq++;
	}
	h->launch_seq++;
	h->tiles_done += (unsigned)kp.ntiles;
	h->last_was_launch = !h->h_mirror;	/* a mirror copy behind the launch is an ordinary stream operation */
	h->st.kernel_launches++;
	h->st.grid = grid;
	h->rows_done += nrows;
	h->st.samples_done += (uint64_t) nrows * h->row_samples;
	return 0;
}

/* after new samples were placed behind the carried tail in the staging buffer */
static int run_staged(vdl2gpu * h, size_t total)
{
	const int nrows = (int)(total / h->row_samples);
	if (run_rows(h, h->d_stage, h->stage_pitch, nrows))
		return 1;
	const size_t used = (size_t) nrows * h->row_samples;
	const size_t tail = total - used;
	if (tail && used) {	/* tail < one row <= used: source and destination never overlap */
		CK(h, cudaMemcpy2DAsync(h->d_stage, h->stage_pitch, h->d_stage + used * h->bytes_per_sample, h->stage_pitch,
					tail * h->bytes_per_sample, h->nstreams, cudaMemcpyDeviceToDevice, h->stream));
	}
	h->carry = tail;
	return 0;
}

extern "C" int vdl2_submit_host(vdl2gpu_t * h, const void *iq, size_t nsamples, size_t pitch_bytes)
{
	if (!h || !iq)
		return fail(h, "vdl2_submit_host: null argument");
	CK(h, cudaSetDevice(h->cfg.device));
	const size_t bps = h->bytes_per_sample;
	if ((h->carry + nsamples) * bps > h->stage_pitch)
		return fail(h, "vdl2_submit_host: %zu samples exceed max_samples of the handle", nsamples);
	if (h->nstreams > 1 && pitch_bytes < nsamples * bps)
		return fail(h, "vdl2_submit_host: pitch %zu smaller than a stream (%zu bytes)", pitch_bytes, nsamples * bps);
	h->st.samples_in += nsamples;
	if (nsamples)
		CK(h, cudaMemcpy2DAsync(h->d_stage + h->carry * bps, h->stage_pitch, iq, h->nstreams > 1 ? pitch_bytes : nsamples * bps,
					nsamples * bps, h->nstreams, cudaMemcpyHostToDevice, h->stream));
	return run_staged(h, h->carry + nsamples);
}

extern "C" int vdl2_process_host(vdl2gpu_t * h, const void *iq, size_t nsamples, size_t pitch_bytes)
{
	if (vdl2_submit_host(h, iq, nsamples, pitch_bytes))
		return 1;
	CK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

/* the caller's buffer is copied into a page-locked ring slot first, so it may be reused as soon as the call returns (the
   reference's Cbuff is overwritten by the next SDR callback, rtl.c:283-294) and the upload runs at the PCIe rate */
extern "C" int vdl2_submit_copy(vdl2gpu_t * h, const void *iq, size_t nsamples, size_t pitch_bytes)
{
	if (!h || !iq)
		return fail(h, "vdl2_submit_copy: null argument");
	CK(h, cudaSetDevice(h->cfg.device));
	const size_t bps = h->bytes_per_sample, row = nsamples * bps;
	if (h->nstreams > 1 && pitch_bytes < row)
		return fail(h, "vdl2_submit_copy: pitch %zu smaller than a stream (%zu bytes)", pitch_bytes, row);
	const size_t need = row * h->nstreams;
	if (need > h->pin_bytes) {	/* first use, or a larger call than before: (re)build the ring */
		CK(h, cudaStreamSynchronize(h->stream));
		for (int i = 0; i < VDL2_PIN_SLOTS; i++) {
			if (h->pin[i])
				cudaFreeHost(h->pin[i]);
			h->pin[i] = NULL;
		}
		h->pin_bytes = 0;
		for (int i = 0; i < VDL2_PIN_SLOTS; i++) {
			CK(h, cudaHostAlloc((void **)&h->pin[i], need, cudaHostAllocPortable));
			if (!h->pin_ev[i])
				CK(h, cudaEventCreateWithFlags(&h->pin_ev[i], cudaEventDisableTiming));
		}
		h->pin_bytes = need;
	}
	const unsigned slot = h->pin_next++ % VDL2_PIN_SLOTS;
	CK(h, cudaEventSynchronize(h->pin_ev[slot]));	/* the upload that last used this slot (a never-recorded event is complete) */
	for (int s_ = 0; s_ < h->nstreams; s_++)
		memcpy(h->pin[slot] + (size_t) s_ * row, (const uint8_t *)iq + (size_t) s_ * pitch_bytes, row);
	if (vdl2_submit_host(h, h->pin[slot], nsamples, row))
		return 1;
	CK(h, cudaEventRecord(h->pin_ev[slot], h->stream));
	return 0;
}

/* blocks completed by launches that have FINISHED, without waiting for one that is still running (may lag by one launch) */
extern "C" int vdl2_pending_blocks(vdl2gpu_t * h, int *n_out)
{
	if (!h || !n_out)
		return fail(h, "vdl2_pending_blocks: null argument");
	*n_out = 0;
	CK(h, cudaSetDevice(h->cfg.device));
	if (!h->h_mirror) {	/* first poll: from now on every launch refreshes the mirror */
		CK(h, cudaHostAlloc((void **)&h->h_mirror, VDL2_MIRROR_SLOTS * 64, cudaHostAllocPortable));
		memset(h->h_mirror, 0, VDL2_MIRROR_SLOTS * 64);
		for (int i = 0; i < VDL2_MIRROR_SLOTS; i++)
			CK(h, cudaEventCreateWithFlags(&h->mirror_ev[i], cudaEventDisableTiming));
		CK(h, cudaMemcpyAsync(h->h_mirror, h->d_ticket, 64, cudaMemcpyDeviceToHost, h->stream));
		CK(h, cudaEventRecord(h->mirror_ev[0], h->stream));
		h->mirror_seq = 1;
	}
	/* newest mirror whose copy has completed */
	for (unsigned back = 1; back <= VDL2_MIRROR_SLOTS && back <= h->mirror_seq; back++) {
		const unsigned q = (h->mirror_seq - back) % VDL2_MIRROR_SLOTS;
		const cudaError_t e = cudaEventQuery(h->mirror_ev[q]);
		if (e == cudaSuccess) {
			*n_out = (int)std::min(h->h_mirror[16 * q + 4], h->outq_cap);
			return 0;
		}
		if (e != cudaErrorNotReady)
			return fail(h, "vdl2_pending_blocks: %s", cudaGetErrorString(e));
	}
	return 0;
}

/* rtl.c:285-292 on the device: out[k] of every 32768-sample block = 0 for k == 0, else (u8 - 127.37f) of sample k - 1.
   10 bytes of HBM traffic per sample, a few microseconds per launch next to the demodulator. */
#define RTL_BLOCK 32768u
__global__ void vdl2_expand_rtl_kernel(const uchar2 * __restrict__ raw, float2 * __restrict__ out, size_t n)
{
	for (size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t) gridDim.x * blockDim.x) {
		float2 v = make_float2(0.0f, 0.0f);
		if ((unsigned)(k & (RTL_BLOCK - 1)) != 0u) {
			const uchar2 u = raw[k - 1];
			v.x = __fsub_rn((float)u.x, 127.37f);	/* (float)x - (float)127.37, rtl.c:287-289 */
			v.y = __fsub_rn((float)u.y, 127.37f);
		}
		out[k] = v;
	}
}

extern "C" int vdl2_process_host_rtl(vdl2gpu_t * h, const void *cu8, size_t nsamples)
{
	if (!h || !cu8)
		return fail(h, "vdl2_process_host_rtl: null argument");
	if (h->cfg.format != VDL2_FMT_CF32 || h->nstreams != 1)
		return fail(h, "vdl2_process_host_rtl: needs a handle with format VDL2_FMT_CF32 and one stream");
	if (nsamples % RTL_BLOCK)
		return fail(h, "vdl2_process_host_rtl: %zu samples are not whole %u-sample callbacks", nsamples, RTL_BLOCK);
	CK(h, cudaSetDevice(h->cfg.device));
	const size_t bps = h->bytes_per_sample;	/* 8: the staging buffer holds complex float */
	if ((h->carry + nsamples) * bps > h->stage_pitch)
		return fail(h, "vdl2_process_host_rtl: %zu samples exceed max_samples of the handle", nsamples);
	h->st.samples_in += nsamples;
	if (nsamples) {
		if (2 * nsamples > h->raw_cap) {
			CK(h, cudaStreamSynchronize(h->stream));
			cudaFree(h->d_raw);
			h->d_raw = NULL;
			h->raw_cap = 0;
			CK(h, cudaMalloc(&h->d_raw, 2 * nsamples));
			h->raw_cap = 2 * nsamples;
		}
		CK(h, cudaMemcpyAsync(h->d_raw, cu8, 2 * nsamples, cudaMemcpyHostToDevice, h->stream));
		const int threads = 256;
		const int blocks = (int)std::min < size_t > ((nsamples + threads - 1) / threads, (size_t) h->st.n_sm * 8);
		vdl2_expand_rtl_kernel <<< blocks, threads, 0, h->stream >>> ((const uchar2 *)h->d_raw, (float2 *) (h->d_stage + h->carry * bps), nsamples);
		CK(h, cudaGetLastError());
	}
	if (run_staged(h, h->carry + nsamples))
		return 1;
	CK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

extern "C" int vdl2_process_device(vdl2gpu_t * h, const void *d_iq, size_t nsamples, size_t pitch_bytes)
{
	if (!h || !d_iq)
		return fail(h, "vdl2_process_device: null argument");
	CK(h, cudaSetDevice(h->cfg.device));
	const size_t bps = h->bytes_per_sample;
	h->st.samples_in += nsamples;
	const bool aligned = (((uintptr_t) d_iq & 15) == 0) && ((pitch_bytes & 15) == 0 || h->nstreams == 1);
	if (h->carry == 0 && nsamples % h->row_samples == 0 && aligned) {
		/* zero copy: the TMA descriptor points straight at the caller's buffer */
		const int nrows = (int)(nsamples / h->row_samples);
		return run_rows(h, d_iq, h->nstreams == 1 ? (size_t) h->row_bytes * nrows : pitch_bytes, nrows);
	}
	if ((h->carry + nsamples) * bps > h->stage_pitch)
		return fail(h, "vdl2_process_device: %zu samples exceed max_samples of the handle", nsamples);
	if (nsamples)
		CK(h, cudaMemcpy2DAsync(h->d_stage + h->carry * bps, h->stage_pitch, d_iq, h->nstreams > 1 ? pitch_bytes : nsamples * bps,
					nsamples * bps, h->nstreams, cudaMemcpyDeviceToDevice, h->stream));
	return run_staged(h, h->carry + nsamples);
}

/* ---- row f3: wideband shared-stream channeliser (vdl2_channelise_kernel).  One pass over every input stream writes the
   decimated 84 ksps streams (d8psk.c:374-381, tap T1) of ALL the channels demodulated from it; no demodulator state is touched.
   8-bit input at 2 Msps only (the tensor-core mixer's tables); whole 1 ms rows. ---- */
extern "C" int vdl2_channelise_device(vdl2gpu_t * h, const void *d_iq, size_t nsamples, size_t pitch_bytes, float *d_out, size_t out_pitch)
{
	if (!h || !d_iq || !d_out)
		return fail(h, "vdl2_channelise_device: null argument");
	if (h->dp4a != 2)
		return fail(h, "vdl2_channelise_device: needs 8-bit input at 2 Msps (the tensor-core mixer)");
	if (nsamples == 0 || nsamples % h->row_samples)
		return fail(h, "vdl2_channelise_device: %zu samples are not whole 1 ms rows", nsamples);
	const int nrows = (int)(nsamples / h->row_samples);
	if (out_pitch < (size_t) nrows * VDL2_DUMPS_PER_ROW || (out_pitch & 1) || ((uintptr_t) d_out & 15))
		return fail(h, "vdl2_channelise_device: output pitch %zu too small / odd, or output not 16-byte aligned", out_pitch);
	const size_t pitch = h->nstreams == 1 ? (size_t) h->row_bytes * nrows : pitch_bytes;
	if (((uintptr_t) d_iq & 15) || (pitch & 15))
		return fail(h, "vdl2_channelise_device: input base/pitch must be 16-byte aligned for TMA");
	CK(h, cudaSetDevice(h->cfg.device));
	CUtensorMap tmap;
	const cuuint64_t dims[3] = { (cuuint64_t) h->row_samples, (cuuint64_t) nrows, (cuuint64_t) h->nstreams };
	const cuuint64_t strides[2] = { (cuuint64_t) h->row_bytes, (cuuint64_t) pitch };
	const cuuint32_t box[3] = { 32, 32, 1 }, estr[3] = { 1, 1, 1 };
	const CUresult r = ((encode_tiled_t) h->encode_fn) (&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void *)d_iq, dims, strides, box, estr,
							   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
							   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS)
		return fail(h, "cuTensorMapEncodeTiled failed (%d)", (int)r);
	const long long items = (long long)((nrows + VDL2_ROWS_PER_TILE - 1) / VDL2_ROWS_PER_TILE) * h->nstreams;
	const int grid = (int)std::min < long long >(items, (long long)h->n_sm * 16);
	CK(h, cudaMemsetAsync(h->d_ticket + 11, 0, 4, h->stream));
	CK(h, cudaEventRecord(h->ev0, h->stream));
	if (l2_reserve(h, 0))	/* the channeliser keeps nothing on chip and wants all of L2 */
		return -1;
	const cudaError_t e = (cudaError_t) vdl2_channelise_launch(h->cfg.format, &tmap, h->nstreams, h->cfg.ch_per_stream, nrows, h->nbox, h->sched_slot, h->d_w8,
								  h->d_dcorr, d_out, out_pitch, h->d_ticket + 11, grid, h->stream);
	if (e != cudaSuccess)
		return fail(h, "channeliser launch failed: %s", cudaGetErrorString(e));
	CK(h, cudaEventRecord(h->ev1, h->stream));
	h->ev_valid = true;
	h->last_was_launch = false;
	return 0;
}

extern "C" int vdl2_sync(vdl2gpu_t * h)
{
	if (!h)
		return 1;
	CK(h, cudaSetDevice(h->cfg.device));
	CK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

/* Stable sort of large fixed-size records (2 KB blocks and frames): order an index array, then gather once.
   Sorting the records by value moves every one of them about log2(n) times (some 100 MB of copies for the
   4 500 frames of one bench step). */
template < class T, class Less > static void sort_records(T * rec, size_t n, Less less)
{
	if (n < 2)
		return;
	std::vector < uint32_t > idx(n);
	for (size_t i = 0; i < n; i++)
		idx[i] = (uint32_t) i;
	std::stable_sort(idx.begin(), idx.end(),[&](uint32_t a, uint32_t b) {
			 return less(rec[a], rec[b]);}
	);
	size_t first = 0;
	while (first < n && idx[first] == first)
		first++;
	if (first == n)
		return;		/* already in order */
	std::vector < T > tmp(rec + first, rec + n);
	for (size_t i = first; i < n; i++)
		rec[i] = tmp[idx[i] - first];	/* entries before `first` stay put, so every source index is >= first */
}

extern "C" int vdl2_drain_blocks(vdl2gpu_t * h, vdl2_block_t * out, int max, int *n_out)
{
	if (!h || !n_out)
		return fail(h, "vdl2_drain_blocks: null argument");
	*n_out = 0;
	CK(h, cudaSetDevice(h->cfg.device));
	CK(h, cudaStreamSynchronize(h->stream));
	unsigned cnt[8];
	CK(h, cudaMemcpy(cnt, h->d_outq_count, sizeof cnt, cudaMemcpyDeviceToHost));
	unsigned n = std::min(cnt[0], h->outq_cap);
	h->st.blocks_dropped += cnt[4];
	if (n == 0) {
		if (cnt[4])
			CK(h, cudaMemset(h->d_dropped, 0, 4));
		return 0;
	}
	if (!out || max < (int)n)
		return fail(h, "vdl2_drain_blocks: %u blocks pending, room for %d", n, max);
	CK(h, cudaMemcpy(out, h->d_outq, sizeof(Vdl2BlockRec) * n, cudaMemcpyDeviceToHost));
	CK(h, cudaMemset(h->d_outq_count, 0, 4));
	if (cnt[4])
		CK(h, cudaMemset(h->d_dropped, 0, 4));
	if (h->h_mirror)	/* the stream is idle: no mirror copy is in flight */
		for (int i = 0; i < VDL2_MIRROR_SLOTS; i++)
			h->h_mirror[16 * i + 4] = 0;
	sort_records(out, (size_t) n,[](const vdl2_block_t & a, const vdl2_block_t & b) {
		     return a.sync_dump != b.sync_dump ? a.sync_dump < b.sync_dump : a.chn < b.chn;}
	);
	h->st.blocks_out += n;
	*n_out = (int)n;
	return 0;
}

/* ---- ingest: page-locked host buffers (row f2) ---- */
extern "C" int vdl2_host_alloc(size_t bytes, void **out)
{
	if (!out || !bytes)
		return fail(nullptr, "vdl2_host_alloc: null argument");
	*out = nullptr;
	cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
	if (e != cudaSuccess)
		return fail(nullptr, "vdl2_host_alloc: %zu bytes: %s", bytes, cudaGetErrorString(e));
	return 0;
}

extern "C" int vdl2_host_free(void *p)
{
	if (!p)
		return 0;
	cudaError_t e = cudaFreeHost(p);
	if (e != cudaSuccess)
		return fail(nullptr, "vdl2_host_free: %s", cudaGetErrorString(e));
	return 0;
}

/* ---- block pipeline ---- */
static int link_reserve(vdl2gpu * h, int nblocks, int nframes, bool own_blocks)
{
	if (!h->link_ready) {
		cudaError_t e = (cudaError_t) vdl2_link_upload_tables();
		if (e != cudaSuccess)
			return fail(h, "block pipeline: table upload failed: %s", cudaGetErrorString(e));
		CK(h, cudaEventCreate(&h->lev0));
		CK(h, cudaEventCreate(&h->lev1));
		CK(h, cudaMalloc(&h->d_nframes, 16));
		h->link_ready = true;
	}
	if (nblocks > h->lcap_blocks) {
		cudaFree(h->d_lblocks);
		cudaFree(h->d_lstats);
		cudaFree(h->d_lrows);
		h->d_lblocks = NULL;
		h->d_lstats = NULL;
		h->d_lrows = NULL;
		const int cap = std::max(nblocks, 256);
		if (own_blocks)
			CK(h, cudaMalloc(&h->d_lblocks, sizeof(Vdl2BlockRec) * (size_t) cap));
		CK(h, cudaMalloc(&h->d_lstats, sizeof(Vdl2BlkStat) * (size_t) cap));
		CK(h, cudaMalloc(&h->d_lrows, (size_t) 2040 * cap));
		h->lcap_blocks = cap;
	} else if (own_blocks && !h->d_lblocks) {
		CK(h, cudaMalloc(&h->d_lblocks, sizeof(Vdl2BlockRec) * (size_t) h->lcap_blocks));
	}
	if (nframes > h->lcap_frames) {
		cudaFree(h->d_frames);
		h->d_frames = NULL;
		const int cap = std::max(nframes, 256);
		CK(h, cudaMalloc(&h->d_frames, sizeof(Vdl2FrameRec) * (size_t) cap));
		h->lcap_frames = cap;
	}
	return 0;
}

/* runs the kernel on nblocks device-resident blocks; copies the frames back, sorted by (block, length) */
static int link_run(vdl2gpu * h, const Vdl2BlockRec * d_blocks, int nblocks, vdl2_frame_t * frames, int max_frames, int *n_frames,
		    bool want_rows)
{
	*n_frames = 0;
	if (nblocks <= 0)
		return 0;
	CK(h, cudaMemsetAsync(h->d_nframes, 0, 8, h->stream));
	CK(h, cudaEventRecord(h->lev0, h->stream));
	cudaError_t e = (cudaError_t) vdl2_link_launch(d_blocks, nblocks, h->d_frames, h->d_nframes, (unsigned)std::min(max_frames, h->lcap_frames),
						       h->d_lstats, want_rows ? h->d_lrows : NULL, h->stream);
	if (e != cudaSuccess)
		return fail(h, "block pipeline launch failed: %s", cudaGetErrorString(e));
	CK(h, cudaEventRecord(h->lev1, h->stream));
	h->lev_valid = true;
	h->st.link_launches++;
	unsigned nfv[2] = { 0, 0 };
	CK(h, cudaMemcpyAsync(nfv, h->d_nframes, 8, cudaMemcpyDeviceToHost, h->stream));
	CK(h, cudaStreamSynchronize(h->stream));
	const unsigned nf = nfv[0];
	if (nfv[1])	/* more FCS-good candidates in one block than the kernel hands over: never silently */
		return fail(h, "block pipeline: %u candidate frames beyond the per-block limit were not delivered", nfv[1]);
	if ((int)nf > max_frames || (int)nf > h->lcap_frames)
		return fail(h, "block pipeline: %u frames, room for %d", nf, std::min(max_frames, h->lcap_frames));
	if (nf) {
		CK(h, cudaMemcpy(frames, h->d_frames, sizeof(Vdl2FrameRec) * (size_t) nf, cudaMemcpyDeviceToHost));
		std::vector < int >idx(nf);
		for (unsigned i = 0; i < nf; i++)
			idx[i] = (int)i;
		std::sort(idx.begin(), idx.end(),[&](int a, int b) {
			  return frames[a].block != frames[b].block ? frames[a].block < frames[b].block : frames[a].len < frames[b].len;}
		);
		std::vector < vdl2_frame_t > tmp(frames, frames + nf);
		for (unsigned i = 0; i < nf; i++)
			frames[i] = tmp[idx[i]];
	}
	h->st.frames_out += nf;
	*n_frames = (int)nf;
	return 0;
}

extern "C" int vdl2_link_decode(vdl2gpu_t * h, const vdl2_block_t * blocks, int nblocks, vdl2_frame_t * frames, int max_frames, int *n_frames,
				vdl2_blkstat_t * stats, uint8_t * rows_after)
{
	if (!h || !n_frames || (nblocks > 0 && (!blocks || !frames)))
		return fail(h, "vdl2_link_decode: null argument");
	*n_frames = 0;
	if (nblocks <= 0)
		return 0;
	CK(h, cudaSetDevice(h->cfg.device));
	if (link_reserve(h, nblocks, max_frames, true))
		return 1;
	CK(h, cudaMemcpyAsync(h->d_lblocks, blocks, sizeof(Vdl2BlockRec) * (size_t) nblocks, cudaMemcpyHostToDevice, h->stream));
	if (link_run(h, h->d_lblocks, nblocks, frames, max_frames, n_frames, rows_after != NULL))
		return 1;
	if (stats)
		CK(h, cudaMemcpy(stats, h->d_lstats, sizeof(Vdl2BlkStat) * (size_t) nblocks, cudaMemcpyDeviceToHost));
	if (rows_after)
		CK(h, cudaMemcpy(rows_after, h->d_lrows, (size_t) 2040 * nblocks, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int vdl2_drain_frames(vdl2gpu_t * h, vdl2_frame_t * frames, int max_frames, int *n_frames, vdl2_block_t * blocks, int max_blocks,
				 int *n_blocks)
{
	if (!h || !n_frames || !frames)
		return fail(h, "vdl2_drain_frames: null argument");
	*n_frames = 0;
	if (n_blocks)
		*n_blocks = 0;
	CK(h, cudaSetDevice(h->cfg.device));
	CK(h, cudaStreamSynchronize(h->stream));
	unsigned cnt[8];
	CK(h, cudaMemcpy(cnt, h->d_outq_count, sizeof cnt, cudaMemcpyDeviceToHost));
	const unsigned n = std::min(cnt[0], h->outq_cap);
	h->st.blocks_dropped += cnt[4];
	if (cnt[4])
		CK(h, cudaMemset(h->d_dropped, 0, 4));
	if (n == 0)
		return 0;
	if (blocks && max_blocks < (int)n)
		return fail(h, "vdl2_drain_frames: %u blocks pending, room for %d", n, max_blocks);
	if (link_reserve(h, (int)n, max_frames, false))
		return 1;
	/* the block queue is in completion order; the kernel indexes it as it is, the host sorts afterwards */
	if (link_run(h, h->d_outq, (int)n, frames, max_frames, n_frames, false))
		return 1;
	CK(h, cudaMemsetAsync(h->d_outq_count, 0, 4, h->stream));
	if (h->h_mirror)
		for (int i = 0; i < VDL2_MIRROR_SLOTS; i++)
			h->h_mirror[16 * i + 4] = 0;
	if (blocks) {
		/* order of the blocks: oldest trigger first (the order drain_blocks returns); frame.block follows it */
		std::vector < vdl2_block_t > q(n);
		CK(h, cudaMemcpy(q.data(), h->d_outq, sizeof(Vdl2BlockRec) * (size_t) n, cudaMemcpyDeviceToHost));
		std::vector < int >order(n), rank(n);
		for (unsigned i = 0; i < n; i++)
			order[i] = (int)i;
		std::stable_sort(order.begin(), order.end(),[&](int a, int b) {
				 return q[a].sync_dump != q[b].sync_dump ? q[a].sync_dump < q[b].sync_dump : q[a].chn < q[b].chn;}
		);
		for (unsigned i = 0; i < n; i++)
			rank[order[i]] = (int)i;
		for (int i = 0; i < *n_frames; i++)
			frames[i].block = rank[frames[i].block];
		for (unsigned i = 0; i < n; i++)
			blocks[i] = q[order[i]];
	} else {
		for (int i = 0; i < *n_frames; i++)
			frames[i].block = -1;	/* the blocks were not asked for; chn / Fr / ppm / sync_dump travel in the frame */
	}
	/* same order either way: by trigger time, then channel, then position inside the block */
	sort_records(frames, (size_t) * n_frames,[](const vdl2_frame_t & a, const vdl2_frame_t & b) {
		     return a.sync_dump != b.sync_dump ? a.sync_dump < b.sync_dump : (a.chn != b.chn ? a.chn < b.chn : a.len < b.len);}
	);
	h->st.blocks_out += n;
	if (n_blocks)
		*n_blocks = (int)n;
	return 0;
}

/* ---- rows f1 + f4 end to end: the completed blocks never leave the device as blocks.  Block pipeline -> frames, ranked in
   completion order, packed (32-byte header + the frame's own bytes, 16-byte aligned) with their field records, three small
   asynchronous copies.  The fixed 2048-byte records of vdl2_drain_frames() cost 9 MB of synchronous pageable copy and a host
   sort per bench step (26-52 ms); this is a few hundred microseconds. ---- */
extern "C" int vdl2_drain_frames_packed(vdl2gpu_t * h, vdl2_frame_hdr_t * hdrs, int max_frames, int *n_frames, uint8_t * bytes, size_t max_bytes,
					size_t *n_bytes, vdl2_avlc_t * recs)
{
	if (!h || !n_frames || !n_bytes || !hdrs || !bytes)
		return fail(h, "vdl2_drain_frames_packed: null argument");
	static_assert(sizeof(vdl2_frame_hdr_t) == 32, "frame header layout");
	*n_frames = 0;
	*n_bytes = 0;
	CK(h, cudaSetDevice(h->cfg.device));
	CK(h, cudaStreamSynchronize(h->stream));
	unsigned cnt[8];
	CK(h, cudaMemcpy(cnt, h->d_outq_count, sizeof cnt, cudaMemcpyDeviceToHost));
	const unsigned n = std::min(cnt[0], h->outq_cap);
	h->st.blocks_dropped += cnt[4];
	if (cnt[4])
		CK(h, cudaMemset(h->d_dropped, 0, 4));
	if (n == 0)
		return 0;
	const int fcap = std::max(max_frames, 256);
	if (link_reserve(h, (int)n, fcap, false))
		return 1;
	if (fcap > h->pack_cap) {
		cudaFree(h->d_rank);
		cudaFree(h->d_offs);
		cudaFree(h->d_hdrs);
		cudaFree(h->d_precs);
		cudaFree(h->d_keys);
		cudaFree(h->d_pbytes);
		h->d_keys = NULL;
		h->d_rank = NULL;
		h->d_offs = NULL;
		h->d_hdrs = h->d_precs = NULL;
		h->d_pbytes = NULL;
		h->pack_cap = 0;
		CK(h, cudaMalloc(&h->d_rank, sizeof(int) * (size_t) fcap));
		CK(h, cudaMalloc(&h->d_offs, sizeof(unsigned) * (size_t) fcap));
		CK(h, cudaMalloc(&h->d_hdrs, 32 * (size_t) fcap));
		CK(h, cudaMalloc(&h->d_precs, sizeof(vdl2_avlc_t) * (size_t) fcap));
		CK(h, cudaMalloc(&h->d_keys, 16 * (size_t) fcap));
		CK(h, cudaMalloc(&h->d_pbytes, (size_t) 2032 * fcap));	/* worst case: every frame 2016 bytes + padding */
		h->pack_cap = fcap;
	}
	if (!h->d_totals) {
		CK(h, cudaMalloc(&h->d_totals, 16));
		CK(h, cudaHostAlloc((void **)&h->h_totals, 16, cudaHostAllocPortable));
		CK(h, cudaEventCreate(&h->pev0));
		CK(h, cudaEventCreate(&h->pev1));
	}
	const unsigned cap = (unsigned)std::min(fcap, h->lcap_frames);
	CK(h, cudaMemsetAsync(h->d_nframes, 0, 8, h->stream));
	CK(h, cudaEventRecord(h->lev0, h->stream));
	cudaError_t e = (cudaError_t) vdl2_link_launch(h->d_outq, (int)n, h->d_frames, h->d_nframes, cap, h->d_lstats, NULL, h->stream);
	if (e != cudaSuccess)
		return fail(h, "block pipeline launch failed: %s", cudaGetErrorString(e));
	CK(h, cudaEventRecord(h->lev1, h->stream));
	h->lev_valid = true;
	h->st.link_launches++;
	CK(h, cudaEventRecord(h->pev0, h->stream));
	e = (cudaError_t) vdl2_frames_pack_launch(h->d_frames, h->d_nframes, cap, h->d_rank, h->d_offs, h->d_totals, h->d_hdrs, h->d_pbytes,
						  (unsigned)std::min < size_t > ((size_t) 2032 * h->pack_cap, 0xffffffffu), recs ? h->d_precs : NULL,
						  (int)std::min < unsigned >(4 * n, cap), h->d_keys, h->stream);
	if (e != cudaSuccess)
		return fail(h, "frame packing launch failed: %s", cudaGetErrorString(e));
	CK(h, cudaEventRecord(h->pev1, h->stream));
	CK(h, cudaMemcpyAsync(h->h_totals, h->d_totals, 8, cudaMemcpyDeviceToHost, h->stream));
	CK(h, cudaMemsetAsync(h->d_outq_count, 0, 4, h->stream));
	CK(h, cudaStreamSynchronize(h->stream));
	if (h->h_mirror)
		for (int i = 0; i < VDL2_MIRROR_SLOTS; i++)
			h->h_mirror[16 * i + 4] = 0;
	h->st.blocks_out += n;
	const unsigned nf = h->h_totals[0], nb = h->h_totals[1];
	unsigned rawv[2] = { 0, 0 };
	CK(h, cudaMemcpy(rawv, h->d_nframes, 8, cudaMemcpyDeviceToHost));
	const unsigned raw = rawv[0];
	if (rawv[1])
		return fail(h, "block pipeline: %u candidate frames beyond the per-block limit were not delivered", rawv[1]);
	if (raw > cap || (int)nf > max_frames)
		return fail(h, "vdl2_drain_frames_packed: %u frames, room for %d", raw, std::min(max_frames, (int)cap));
	if (nb > max_bytes)
		return fail(h, "vdl2_drain_frames_packed: %u bytes of frames, room for %zu", nb, max_bytes);
	if (nf) {
		CK(h, cudaMemcpyAsync(hdrs, h->d_hdrs, 32 * (size_t) nf, cudaMemcpyDeviceToHost, h->stream));
		CK(h, cudaMemcpyAsync(bytes, h->d_pbytes, nb, cudaMemcpyDeviceToHost, h->stream));
		if (recs)
			CK(h, cudaMemcpyAsync(recs, h->d_precs, sizeof(vdl2_avlc_t) * (size_t) nf, cudaMemcpyDeviceToHost, h->stream));
		CK(h, cudaStreamSynchronize(h->stream));
	}
	{
		float ms = 0;
		if (cudaEventElapsedTime(&ms, h->pev0, h->pev1) == cudaSuccess)
			h->last_pack_ms = ms;
	}
	h->st.frames_out += nf;
	*n_frames = (int)nf;
	*n_bytes = nb;
	return 0;
}

/* device time of the ranking + packing + field kernels of the last vdl2_drain_frames_packed() (CUDA events) */
extern "C" float vdl2_last_pack_ms(const vdl2gpu_t * h)
{
	return h ? h->last_pack_ms : 0.f;
}

/* ---- frame fields (row f4) ---- */
extern "C" int vdl2_avlc_extract(vdl2gpu_t * h, const vdl2_frame_t * frames, int nframes, vdl2_avlc_t * recs)
{
	if (!h || (nframes > 0 && (!frames || !recs)))
		return fail(h, "vdl2_avlc_extract: null argument");
	if (nframes <= 0)
		return 0;
	static_assert(sizeof(vdl2_frame_t) == sizeof(Vdl2FrameRec) && sizeof(vdl2_avlc_t) == 48, "record layouts");
	CK(h, cudaSetDevice(h->cfg.device));
	if (link_reserve(h, 0, nframes, false))
		return 1;
	if (nframes > h->avlc_cap) {
		cudaFree(h->d_avlc);
		h->d_avlc = NULL;
		h->avlc_cap = 0;
		const int cap = std::max(nframes, 256);
		CK(h, cudaMalloc(&h->d_avlc, sizeof(vdl2_avlc_t) * (size_t) cap));
		h->avlc_cap = cap;
	}
	CK(h, cudaMemcpyAsync(h->d_frames, frames, sizeof(Vdl2FrameRec) * (size_t) nframes, cudaMemcpyHostToDevice, h->stream));
	cudaError_t e = (cudaError_t) vdl2_avlc_launch(h->d_frames, nframes, h->d_avlc, h->stream);
	if (e != cudaSuccess)
		return fail(h, "frame field kernel launch failed: %s", cudaGetErrorString(e));
	CK(h, cudaMemcpyAsync(recs, h->d_avlc, sizeof(vdl2_avlc_t) * (size_t) nframes, cudaMemcpyDeviceToHost, h->stream));
	CK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

template < class T > static int read_tap(vdl2gpu * h, int ch, T * d_base, unsigned cap, size_t off_count, T * out, size_t max,
					  size_t *n_out)
{
	if (!h || !n_out)
		return fail(h, "vdl2_read_*: null argument");
	*n_out = 0;
	if (!d_base)
		return fail(h, "vdl2_read_*: this tap was not enabled in vdl2_config_t.taps");
	if (ch < 0 || ch >= h->cfg.nch)
		return fail(h, "vdl2_read_*: channel %d out of range", ch);
	CK(h, cudaSetDevice(h->cfg.device));
	CK(h, cudaStreamSynchronize(h->stream));
	unsigned n = 0;
	char *cnt = (char *)(h->d_state + ch) + off_count;
	CK(h, cudaMemcpy(&n, cnt, 4, cudaMemcpyDeviceToHost));
	if (n > cap)
		return fail(h, "vdl2_read_*: tap overflow on channel %d (%u records, capacity %u); read more often", ch, n, cap);
	if (n > max)
		return fail(h, "vdl2_read_*: %u records pending, room for %zu", n, max);
	if (n)
		CK(h, cudaMemcpy(out, d_base + (size_t) ch * cap, sizeof(T) * n, cudaMemcpyDeviceToHost));
	CK(h, cudaMemset(cnt, 0, 4));
	*n_out = n;
	return 0;
}

extern "C" int vdl2_read_dumps(vdl2gpu_t * h, int ch, float *iq_out, size_t max, size_t *n_out)
{
	return read_tap < float2 > (h, ch, h ? h->d_tap_dumps : NULL, h ? h->cap_dumps : 0, offsetof(Vdl2ChanState, n_dumps),
				    (float2 *) iq_out, max, n_out);
}

extern "C" int vdl2_read_steps(vdl2gpu_t * h, int ch, vdl2_step_t * out, size_t max, size_t *n_out)
{
	return read_tap < Vdl2StepRec > (h, ch, h ? h->d_tap_steps : NULL, h ? h->cap_steps : 0, offsetof(Vdl2ChanState, n_steps),
					 (Vdl2StepRec *) out, max, n_out);
}

extern "C" int vdl2_read_syncs(vdl2gpu_t * h, int ch, vdl2_sync_t * out, size_t max, size_t *n_out)
{
	return read_tap < Vdl2SyncRec > (h, ch, h ? h->d_tap_syncs : NULL, h ? h->cap_syncs : 0, offsetof(Vdl2ChanState, n_syncs),
					 (Vdl2SyncRec *) out, max, n_out);
}

extern "C" int vdl2_read_syms(vdl2gpu_t * h, int ch, vdl2_sym_t * out, size_t max, size_t *n_out)
{
	return read_tap < Vdl2SymRec > (h, ch, h ? h->d_tap_syms : NULL, h ? h->cap_syms : 0, offsetof(Vdl2ChanState, n_syms),
					(Vdl2SymRec *) out, max, n_out);
}

extern "C" int vdl2_get_stats(vdl2gpu_t * h, vdl2_stats_t * st)
{
	if (!h || !st)
		return 1;
	CK(h, cudaSetDevice(h->cfg.device));
	if (h->ev_valid) {
		CK(h, cudaEventSynchronize(h->ev1));
		float ms = 0;
		CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
		h->st.last_kernel_ms = ms;
	}
	if (h->lev_valid) {
		CK(h, cudaEventSynchronize(h->lev1));
		float ms = 0;
		CK(h, cudaEventElapsedTime(&ms, h->lev0, h->lev1));
		h->st.last_link_ms = ms;
	}
	*st = h->st;
	return 0;
}

extern "C" void *vdl2_cuda_stream(vdl2gpu_t * h)
{
	return h ? (void *)h->stream : NULL;
}
