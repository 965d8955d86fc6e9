/*
 * tests/emul/emul_main.cpp -- TEST INFRASTRUCTURE: runs phase 2 of the kernel
 * (vdl2_demod.cuh, unmodified) on the host through the fibre warp emulator.
 * Input is a decimated 84 ksps stream (e.g. the oracle's tap T1); output are the same
 * block / sync / symbol / step records the kernel writes.
 */
#include <ucontext.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "vdl2_demod.cuh"

namespace vw {
int g_lane;
uint64_t g_xch[32];
static ucontext_t g_main, g_ctx[32];
static int g_arrived, g_gen, g_done[32];
static void yield_() { swapcontext(&g_ctx[g_lane], &g_main); }
void barrier()
{
	const int my = g_gen;
	if (++g_arrived == 32) {
		g_arrived = 0;
		g_gen++;
		return;
	}
	while (g_gen == my)
		yield_();
}
}

struct TileJob {
	const Vdl2KParams *kp;
	int ch, chn, Fr, nd;
	long long dump_base;
	float2 *sd;
	vdl2::IdleScratch S;
	float *hv;
	int nph[32];
	int pre_mode;		/* 0: no speculative pass A; 1: with the right clock guess; 2: with a wrong / stale guess */
	float2 hist[VDL2_HIST];	/* the real history, made visible only after the prepass (as in the kernel) */
	float ph_hist[VDL2_PHHIST];
	vdl2::ChanRegs R0;
	vdl2::ChanRegs Rout[32];
};
static TileJob *g_job;
static long g_bp_symbols[2];	/* symbols whose phase was computed ahead of the chain: on the right grid, on a wrong one */
extern "C" long emul_burst_pre_symbols(int wrong) { return g_bp_symbols[wrong ? 1 : 0]; }

static void lane_main(int lane)
{
	TileJob *j = g_job;
	vdl2::ChanRegs R = j->R0;
	int nph = 0;
	vdl2::IdlePre pre;
	pre.valid = 0;
	pre.used = 0;
	pre.pos0 = 0;
	vdl2::BurstPre bp;
	bp.valid = 0;
	if (j->pre_mode) {
		/* the kernel runs the speculative stage BEFORE it has the previous tile's state: only a guess of (state, clk) */
		const int cg = j->pre_mode == 1 ? R.clk : ((R.clk + 3) & 7);
		const int sg = j->pre_mode == 1 ? R.state : VDL2_ST_WSYNC;
		const long tix = (long)(j->dump_base / (j->nd > 0 ? j->nd : 1));
		if (R.state == VDL2_ST_GETDATA && (j->pre_mode == 1 || (tix & 1))) {
			/* a tile that starts inside a burst whose header is known: phases ahead (BurstPre), mode 1 on the true symbol grid and
			   tap phase, mode 2 on a wrong one (must be ignored) */
			const int c = R.clk;
			const int k0 = (c >= 28) ? 1 : ((35 - c) >> 2);
			const int r = c + 4 * k0 - 32, ds0 = k0 - 1;
			if (r >= 0 && r < 4) {
				const vdl2::BurstGeom g = vdl2::burst_geom(R.nbrow, R.nlbyte);
				const int end = ds0 + 8 * (g.nsym - R.symidx - 1);	/* tile dump of the last symbol of the burst */
				const int last = end < j->nd - 1 ? end : j->nd - 1;
				const int wrong = j->pre_mode == 2;
				const int d0 = (ds0 + ((wrong && (tix & 2)) ? 3 : 0)) & 7;
				const int rr = (wrong && !(tix & 2)) ? (r + 1) & 3 : r;
				if (last >= 24) {
					vdl2::burst_prephase(*j->kp, j->sd, j->S, bp, d0, last, rr, (wrong || !(tix & 4)) ? R.df : R.df + 1e-3f);
					if (lane == 0)
						g_bp_symbols[wrong] += (last - bp.d0) / 8 + 1;
				}
				if (end < j->nd - 1 && !(wrong && (tix & 2))) {
					/* the burst ends inside the tile: speculative idle search of the rest (IdlePre.pos0), with the right tick clock or,
					   in mode 2, a wrong one */
					vdl2::ChanRegs G;
					memset(&G, 0, sizeof G);
					G.clk = wrong ? (r + 2) & 3 : r;
					G.state = VDL2_ST_WSYNC;
					G.perr = 100.f;
					int nph0 = 0;
					pre.pos0 = end + 1;
					vdl2::demod_tile < true > (*j->kp, j->ch, j->chn, j->Fr, G, j->sd, j->S, j->hv, j->nd, j->dump_base, nph0, pre, true, bp);
					if (lane == 0 && pre.valid)
						g_bp_symbols[wrong] += 1000000;
				}
			}
		} else
		if (sg == VDL2_ST_WSYNC && cg >= 0 && cg < 8) {
			vdl2::ChanRegs G;
			memset(&G, 0, sizeof G);
			G.clk = cg;
			G.state = VDL2_ST_WSYNC;
			G.perr = 100.f;
			int nph0 = 0;
			pre.pos0 = 0;
			vdl2::demod_tile < true > (*j->kp, j->ch, j->chn, j->Fr, G, j->sd, j->S, j->hv, j->nd, j->dump_base, nph0, pre, true, bp);
		}
		vw::sync();
		if (lane < VDL2_HIST)
			j->sd[lane] = j->hist[lane];
		j->S.pht[lane] = j->ph_hist[lane];
		j->S.pht[lane + 32] = j->ph_hist[lane + 32];
		vw::sync();
	}
	vdl2::demod_tile < true > (*j->kp, j->ch, j->chn, j->Fr, R, j->sd, j->S, j->hv, j->nd, j->dump_base, nph, pre, false, bp);
	j->nph[lane] = nph;
	j->Rout[lane] = R;
	vw::g_done[lane] = 1;
}

static void run_warp(TileJob * job)
{
	static std::vector < char >stacks;
	const size_t SS = 256 * 1024;
	if (stacks.empty())
		stacks.resize(32 * SS);
	g_job = job;
	vw::g_arrived = 0;
	for (int i = 0; i < 32; i++) {
		vw::g_done[i] = 0;
		getcontext(&vw::g_ctx[i]);
		vw::g_ctx[i].uc_stack.ss_sp = stacks.data() + i * SS;
		vw::g_ctx[i].uc_stack.ss_size = SS;
		vw::g_ctx[i].uc_link = &vw::g_main;
		makecontext(&vw::g_ctx[i], (void (*)())lane_main, 1, i);
	}
	int live = 32;
	while (live) {
		live = 0;
		for (int i = 0; i < 32; i++) {
			if (vw::g_done[i])
				continue;
			vw::g_lane = i;
			swapcontext(&vw::g_main, &vw::g_ctx[i]);
			if (!vw::g_done[i])
				live++;
		}
	}
}

static bool same_regs(const vdl2::ChanRegs & a, const vdl2::ChanRegs & b)
{
#define EQ(f) (memcmp(&a.f, &b.f, sizeof a.f) == 0)
	return EQ(perr) && EQ(p2err) && EQ(pfr) && EQ(df) && EQ(P1) && EQ(ppm) && EQ(clk) && EQ(state) && EQ(symidx) && EQ(nbrow)
	    && EQ(nlbyte) && EQ(bytes_done) && EQ(bitacc) && EQ(nbitacc) && EQ(sync_dump) && EQ(n_steps) && EQ(n_syncs) && EQ(n_syms);
#undef EQ
}

static void build_tables()
{
	memset(&c_tab, 0, sizeof c_tab);
	float mf[VDL2_MFLTLEN], sw[VDL2_NBPH];
	static float soft[3][257];
	vdl2_make_mflt(mf);
	vdl2_make_sync(sw);
	vdl2_make_softmap(soft);
	memcpy(c_tab.mflt, mf, sizeof mf);
	memcpy(c_tab.sync, sw, sizeof sw);
	for (int b = 0; b < 3; b++)
		memcpy(c_tab.soft[b], soft[b], sizeof soft[b]);
	unsigned s = 0x4D4B;
	for (int i = 0; i < VDL2_SCR_WORDS * 32; i++) {
		unsigned b = (s ^ (s >> 14)) & 1u;
		s = (s << 1) | b;
		c_tab.scr[i >> 5] |= b << (i & 31);
	}
	static const unsigned char hc[25] = { 0x06, 0x07, 0x09, 0x0a, 0x0b, 0x0c, 0x0e, 0x0f, 0x11, 0x13, 0x15, 0x16, 0x18,
		0x19, 0x1a, 0x1b, 0x1c, 0x1d, 0x1e, 0x1f, 0x10, 0x08, 0x04, 0x02, 0x01
	};
	memcpy(c_tab.hcol, hc, sizeof hc);
}

/* Demodulate `ndumps` decimated samples in tiles of `tile_dumps` (<= 2688).  Record buffers
   are caller-owned; counts come back through n_*.  Returns 0. */
extern "C" int emul_demod(const float *dumps, long ndumps, int tile_dumps, int chn, int Fr, unsigned flags, Vdl2BlockRec * blocks, unsigned cap_blocks,
			  unsigned *n_blocks, Vdl2StepRec * steps, unsigned cap_steps, unsigned *n_steps, Vdl2SyncRec * syncs,
			  unsigned cap_syncs, unsigned *n_syncs, Vdl2SymRec * syms, unsigned cap_syms, unsigned *n_syms)
{
	build_tables();
	std::vector < unsigned char >curblk(2048, 0);
	unsigned outq_count = 0, dropped = 0;
	Vdl2KParams kp;
	memset(&kp, 0, sizeof kp);
	kp.nch = 1;
	kp.curblk = curblk.data();
	kp.outq = blocks;
	kp.outq_count = &outq_count;
	kp.outq_cap = cap_blocks;
	kp.dropped = &dropped;
	kp.taps = (steps ? VDL2_TAP_STEPS_BIT : 0u) | VDL2_TAP_SYNCS_BIT | VDL2_TAP_SYMS_BIT;
	kp.flags = flags & 0xffu;
	kp.tap_steps = steps;
	kp.tap_syncs = syncs;
	kp.tap_syms = syms;
	kp.cap_steps = cap_steps;
	kp.cap_syncs = cap_syncs;
	kp.cap_syms = cap_syms;

	std::vector < float2 > sd(VDL2_HIST + VDL2_TILE_DUMPS);
	float hv[32] = { 0 };
	std::vector < float >pht(VDL2_PHT_LEN, 0.f);
	std::vector < float2 > vwin(VDL2_BPRE_BUF);	/* 96 for the idle search; burst_prephase stages its windows in vw .. cand0 (contiguous in the kernel) */
	std::vector < unsigned short >cand(VDL2_CAND_CAP + VDL2_CAND0_CAP);
	std::vector < float2 > win(VDL2_WIN_LEN);
	std::vector < unsigned char >hbuf(VDL2_TILE_DUMPS / 8);
	for (int i = 0; i < VDL2_HIST; i++)
		sd[i] = make_float2(0.f, 0.f);
	vdl2::ChanRegs R;
	memset(&R, 0, sizeof R);
	R.perr = 100.f;
	R.state = VDL2_ST_WSYNC;
	R.sync_dump = -1;
	TileJob job;
	for (long base = 0; base < ndumps; base += tile_dumps) {
		const int nd = (int)((ndumps - base) < tile_dumps ? (ndumps - base) : tile_dumps);
		for (int i = 0; i < nd; i++)
			sd[VDL2_HIST + i] = make_float2(dumps[2 * (base + i)], dumps[2 * (base + i) + 1]);
		job.kp = &kp;
		job.ch = 0;
		job.chn = chn;
		job.Fr = Fr;
		job.nd = nd;
		job.dump_base = base;
		job.sd = sd.data();
		job.S.pht = pht.data();
		job.S.vw = vwin.data();
		job.S.cand = cand.data();
		job.S.cand0 = cand.data() + VDL2_CAND_CAP;
		job.pre_mode = (flags & 0x100u) ? 1 : ((flags & 0x200u) ? 2 : 0);
		if (job.pre_mode) {
			for (int i = 0; i < VDL2_HIST; i++) {
				job.hist[i] = sd[i];
				sd[i] = make_float2(37.f * (float)(i + 1), -11.f * (float)i);	/* stale scratch content */
			}
			for (int i = 0; i < VDL2_PHHIST; i++) {
				job.ph_hist[i] = pht[i];
				pht[i] = 0.1f * (float)i;
			}
		}
		job.S.win = win.data();
		job.S.hb = hbuf.data();
		job.hv = hv;
		job.R0 = R;
		run_warp(&job);
		for (int l = 1; l < 32; l++)
			if (!same_regs(job.Rout[l], job.Rout[0])) {
				fprintf(stderr, "emul: lane %d state diverged from lane 0 at tile base %ld\n", l, base);
				return 2;
			}
		R = job.Rout[0];
		{
			const int nph = job.nph[0];
			for (int i = 0; i < VDL2_PHHIST; i++)
				pht[i] = pht[nph + i];
		}
		for (int i = 0; i < VDL2_HIST; i++)
			sd[i] = sd[nd + i];
	}
	*n_blocks = outq_count;
	*n_steps = R.n_steps;
	*n_syncs = R.n_syncs;
	*n_syms = R.n_syms;
	return 0;
}
