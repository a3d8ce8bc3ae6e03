// capi_clean.cu -- drop-in level C ABI for stage (iii): asg_cut_tip / asg_cut_internal / asg_cut_biloop /
// asg_pop_bubble (asg.h:37-42) and ma_ug_gen (miniasm.h:72).  Host asg_t in, CUDA passes, host structs out.
#include "../../include/miniasm_b200.h"
#include "capi_util.cuh"
#include "clean_dev.cuh"

// unitigs from device buffers into the reference's host structs (ma_ug_t owns malloc'd arrays, asm.c:64-75)
ma_ug_t *mab_ug_download(MabDev &d, DUnitigs &du)
{
	ma_ug_t *ug = (ma_ug_t*)calloc(1, sizeof(ma_ug_t));
	DUtgMeta *meta = (DUtgMeta*)malloc((du.n_utg ? du.n_utg : 1) * sizeof(DUtgMeta));
	uint64_t *items = (uint64_t*)malloc((du.n_items ? du.n_items : 1) * 8);
	if (du.n_utg) MAB_CUDA(cudaMemcpyAsync(meta, du.meta, (size_t)du.n_utg * sizeof(DUtgMeta), cudaMemcpyDeviceToHost, d.stream));
	if (du.n_items) MAB_CUDA(cudaMemcpyAsync(items, du.items, du.n_items * 8, cudaMemcpyDeviceToHost, d.stream));
	d.sync();
	ug->u.n = ug->u.m = du.n_utg;
	ug->u.a = (ma_utg_t*)calloc(du.n_utg ? du.n_utg : 1, sizeof(ma_utg_t));
	for (uint32_t i = 0; i < du.n_utg; ++i) {
		ma_utg_t *p = &ug->u.a[i];
		p->len = meta[i].len, p->circ = meta[i].circ, p->start = meta[i].start, p->end = meta[i].end;
		p->n = meta[i].n, p->m = p->n;
		--p->m; p->m |= p->m >> 1; p->m |= p->m >> 2; p->m |= p->m >> 4; p->m |= p->m >> 8; p->m |= p->m >> 16; ++p->m; // kv_roundup32
		p->a = (uint64_t*)malloc(8 * (size_t)(p->m ? p->m : 1));
		memcpy(p->a, items + meta[i].first, 8 * (size_t)p->n);
		p->s = 0;
	}
	free(meta); free(items);
	asg_t *g = asg_init();
	g->n_seq = du.g.n_seq, g->m_seq = du.g.n_seq ? du.g.n_seq : 1;
	g->seq = (asg_seq_t*)malloc((size_t)g->m_seq * sizeof(asg_seq_t));
	g->m_arc = du.g.n_arc ? du.g.n_arc : 1;
	g->arc = (asg_arc_t*)malloc((size_t)g->m_arc * sizeof(asg_arc_t));
	mab_graph_download(d, du.g, g);
	ug->g = g;
	return ug;
}

extern "C" {

#define WITH_GRAPH(g, body) do { MabDev &d = mab_default_dev(); DGraph dg; mab_graph_upload(d, (g), dg); body; mab_graph_download(d, dg, (g)); dg_free(d, dg); d.sync(); } while (0)

int asg_cut_tip(asg_t *g, int max_ext) { int r; WITH_GRAPH(g, r = (int)dg_cut_tip(d, dg, max_ext)); return r; }
int asg_cut_internal(asg_t *g, int max_ext) { int r; WITH_GRAPH(g, r = (int)dg_cut_internal(d, dg, max_ext)); return r; }
int asg_cut_biloop(asg_t *g, int max_ext) { int r; WITH_GRAPH(g, r = (int)dg_cut_biloop(d, dg, max_ext)); return r; }
int asg_pop_bubble(asg_t *g, int max_dist) { int r; WITH_GRAPH(g, r = (int)dg_pop_bubble(d, dg, max_dist)); return r; }

ma_ug_t *ma_ug_gen(asg_t *g)
{
	MabDev &d = mab_default_dev();
	DGraph dg;
	DUnitigs du;
	mab_graph_upload(d, g, dg);
	if (!dg.has_idx) dg_arc_index(d, dg); // the reference dereferences g->idx unconditionally (asm.c:118-119)
	dg_ug_gen(d, dg, du);
	ma_ug_t *ug = mab_ug_download(d, du);
	dg_ug_free(d, du);
	dg_free(d, dg);
	d.sync();
	return ug;
}

/* rounds / committed actions of the last sequential-semantics pass (tests) */
void mab_last_clean(uint32_t *rounds, uint32_t *committed)
{
	if (rounds) *rounds = g_clean_stats.rounds;
	if (committed) *committed = g_clean_stats.committed;
}

/* all order-dependent passes since the last reset (mab_layout resets): passes run, most sweeps any pass needed, sweeps and actions in total */
void mab_clean_totals(uint32_t *passes, uint32_t *max_sweeps, uint32_t *sweeps, uint32_t *actions)
{
	if (passes) *passes = g_clean_stats.passes;
	if (max_sweeps) *max_sweeps = g_clean_stats.max_rounds;
	if (sweeps) *sweeps = g_clean_stats.sum_rounds;
	if (actions) *actions = g_clean_stats.sum_committed;
}

}
