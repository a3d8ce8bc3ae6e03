// fix_host.cpp -- TEST INFRASTRUCTURE (never linked into the product): the timestamp fixed point of
// miniasm_b200/csrc/clean_fix.cuh compiled for the host and run as plain sequential sweeps, behind the reference's own
// function names (asg.c:238-306,412-433), so that the CPU tier can check the algorithm the CUDA kernels run -- the very
// same header -- against the unmodified reference pass by pass (tests/test_fix_cpu.py).
// The passes here stop after setting the del bits; the caller runs the reference's asg_cleanup when the count is > 0.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <vector>
#include <algorithm>

struct DArc { uint64_t ul; uint32_t v; uint32_t ol_del; };
#include "../../miniasm_b200/csrc/clean_fix.cuh"

extern "C" {
typedef struct { uint32_t len_del; } h_seq_t;
typedef struct {
	uint32_t m_arc, n_arc:31, is_srt:1;
	DArc *arc;
	uint32_t m_seq, n_seq:31, is_symm:1;
	h_seq_t *seq;
	uint64_t *idx;
} h_asg_t;

int fx_last_sweeps = 0, fx_last_nonmono = 0;
}

struct Pass {
	h_asg_t *g;
	std::vector<uint32_t> ts[2], ta[2], is, ia;
	int cur = 0;
	explicit Pass(h_asg_t *g_) : g(g_)
	{
		is.resize(g->n_seq), ia.resize(g->n_arc);
		for (uint32_t i = 0; i < g->n_seq; ++i) is[i] = g->seq[i].len_del & MAB_DEL_BIT ? 0 : FX_LIVE;
		for (uint32_t i = 0; i < g->n_arc; ++i) ia[i] = g->arc[i].ol_del & MAB_DEL_BIT ? 0 : FX_LIVE;
		ts[0] = ts[1] = is, ta[0] = ta[1] = ia;
	}
	FxView view()
	{
		return FxView{g->arc, g->idx, ts[cur].data(), ta[cur].data(), ts[cur ^ 1].data(), ta[cur ^ 1].data(), (uint32_t)g->n_seq * 2};
	}
	bool next() // true if the sweep changed nothing: T_old is the fixed point
	{
		const bool same = ts[0] == ts[1] && ta[0] == ta[1];
		ts[cur] = is, ta[cur] = ia;  // becomes T_new of the next sweep
		cur ^= 1;
		return same;
	}
	void finish() // after next(): cur holds the fixed point
	{
		for (uint32_t i = 0; i < g->n_seq; ++i) if (ts[cur][i] != FX_LIVE) g->seq[i].len_del |= MAB_DEL_BIT;
		for (uint32_t i = 0; i < g->n_arc; ++i) if (ta[cur][i] != FX_LIVE) g->arc[i].ol_del |= MAB_DEL_BIT;
	}
};

template <class Rule> static int run(h_asg_t *g, Rule rule)
{
	fx_last_sweeps = 0;
	if (g->n_seq == 0) return 0;
	{ // probe sweep on the bits (what the device driver does first)
		FxProbe pv{g->arc, g->idx, (const uint32_t*)g->seq, (uint32_t)g->n_seq * 2};
		uint32_t c = 0;
		for (uint32_t x = 0; x < pv.n_vtx; ++x) c += rule.act(pv, x);
		fx_last_sweeps = 1;
		if (c == 0) return 0;
	}
	Pass p(g);
	uint32_t cnt;
	for (;;) {
		FxView v = p.view();
		cnt = 0;
		for (uint32_t x = 0; x < v.n_vtx; ++x) cnt += rule.act(v, x);
		++fx_last_sweeps;
		if (p.next()) break;
	}
	p.finish();
	return (int)cnt;
}

extern "C" {

int asg_cut_tip(h_asg_t *g, int max_ext) { return run(g, FxTip{max_ext}); }
int asg_cut_internal(h_asg_t *g, int max_ext) { return run(g, FxInternal{max_ext}); }
int asg_cut_biloop(h_asg_t *g, int max_ext) { return run(g, FxBiloop{max_ext}); }

// the caller guarantees a symmetric graph (the reference would call asg_symm first)
uint64_t fx_pop_bubble(h_asg_t *g, int max_dist)
{
	fx_last_sweeps = 0, fx_last_nonmono = 0;
	if (g->n_seq == 0 || g->n_arc == 0) return 0;
	Pass p(g);
	uint32_t bcap = 4;  // small on purpose: the grow-and-retry path gets exercised
	uint64_t n_pop, n_tip;
	for (;;) {
		FxView v = p.view();
		bool mono = true, redo;
		do {
			redo = false;
			uint32_t hcap = 1; while (hcap < 2 * bcap) hcap <<= 1;
			std::vector<uint32_t> hkey(hcap, FX_EMPTY), hp(hcap), hd(hcap), hc(hcap), hr(hcap), b(bcap), bs(bcap), S(bcap), e(bcap * 4);
			FxSlot sl{hkey.data(), hp.data(), hd.data(), hc.data(), hr.data(), b.data(), bs.data(), S.data(), e.data(), bcap, bcap * 4, hcap - 1};
			n_pop = n_tip = 0, mono = true;
			for (uint32_t x = 0; x < v.n_vtx && !redo; ++x) {
				uint32_t nt = 0;
				const int r = fx_bub_act(v, x, (uint32_t)max_dist, sl, &nt, &mono);
				if (r < 0) redo = true;
				else if (r == 1) ++n_pop, n_tip += nt;
			}
			if (redo) { // restart this sweep with a larger scratch: T_new back to init
				bcap *= 4;
				p.ts[p.cur ^ 1] = p.is, p.ta[p.cur ^ 1] = p.ia;
			}
		} while (redo);
		++fx_last_sweeps;
		const bool fixed = p.next();
		if (fixed) { fx_last_nonmono = !mono; break; }
	}
	p.finish();
	return (n_pop & 0xffffffffull) | n_tip << 32;
}

}
