// capi_asg.cu -- drop-in level C ABI for the string-graph container (asg.h:31-42 of the reference).
// Each call: host asg_t -> HBM -> CUDA passes (asg_dev.cu / clean_dev.cu) -> host asg_t.
#include "../../include/miniasm_b200.h"
#include "capi_util.cuh"
#include "asg_dev.cuh"

static_assert(sizeof(asg_arc_t) == sizeof(DArc), "arc ABI");
static_assert(sizeof(asg_seq_t) == 4, "seq ABI");

MabDev &mab_default_dev()
{
	static MabDev dev;
	static bool ready = false;
	if (!ready) {
		const char *s = getenv("MINIASM_B200_DEVICE");
		dev.init(s ? atoi(s) : 0);
		ready = true;
	}
	MAB_CUDA(cudaSetDevice(dev.device));
	return dev;
}

void mab_graph_upload(MabDev &d, const asg_t *g, DGraph &dg)
{
	dg_set_nseq(d, dg, g->n_seq);
	dg_reserve(d, dg, g->n_arc ? g->n_arc : 1);
	dg.n_arc = g->n_arc;
	dg.is_srt = g->is_srt, dg.is_symm = g->is_symm;
	if (g->n_arc) MAB_CUDA(cudaMemcpyAsync(dg.arc, g->arc, (size_t)g->n_arc * sizeof(DArc), cudaMemcpyHostToDevice, d.stream));
	if (g->n_seq) MAB_CUDA(cudaMemcpyAsync(dg.seq, g->seq, (size_t)g->n_seq * 4, cudaMemcpyHostToDevice, d.stream));
	if (g->idx && g->n_seq) {
		MAB_CUDA(cudaMemcpyAsync(dg.idx, g->idx, (size_t)g->n_seq * 16, cudaMemcpyHostToDevice, d.stream));
		dg.has_idx = true;
	}
	uint32_t mx = 0;
	for (uint32_t i = 0; i < g->n_arc; ++i) { uint32_t l = (uint32_t)g->arc[i].ul; if (l > mx) mx = l; }
	dg.len_bits = 1; while (dg.len_bits < 32 && (mx >> dg.len_bits)) ++dg.len_bits;
	d.sync();
}

void mab_graph_download(MabDev &d, DGraph &dg, asg_t *g)
{
	if (dg.n_arc > g->m_arc) { // cannot happen for the passes here (they only remove arcs); keep the container sane anyway
		g->m_arc = dg.n_arc;
		g->arc = (asg_arc_t*)realloc(g->arc, (size_t)g->m_arc * sizeof(asg_arc_t));
	}
	if (dg.n_arc) MAB_CUDA(cudaMemcpyAsync(g->arc, dg.arc, (size_t)dg.n_arc * sizeof(DArc), cudaMemcpyDeviceToHost, d.stream));
	if (g->n_seq) MAB_CUDA(cudaMemcpyAsync(g->seq, dg.seq, (size_t)g->n_seq * 4, cudaMemcpyDeviceToHost, d.stream));
	if (g->idx) free(g->idx), g->idx = 0;
	if (dg.has_idx) {
		g->idx = (uint64_t*)calloc((size_t)g->n_seq * 2 + 1, 8);
		if (g->n_seq) MAB_CUDA(cudaMemcpyAsync(g->idx, dg.idx, (size_t)g->n_seq * 16, cudaMemcpyDeviceToHost, d.stream));
	}
	d.sync();
	g->n_arc = dg.n_arc;
	g->is_srt = dg.is_srt, g->is_symm = dg.is_symm;
}

extern "C" {

asg_t *asg_init(void) { return (asg_t*)calloc(1, sizeof(asg_t)); }

void asg_destroy(asg_t *g)
{
	if (g == 0) return;
	free(g->seq); free(g->idx); free(g->arc); free(g);
}

void asg_seq_set(asg_t *g, int sid, int len, int del)
{
	if ((uint32_t)sid >= g->m_seq) {
		uint32_t m = (uint32_t)sid + 1;
		--m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; ++m;
		g->m_seq = m;
		g->seq = (asg_seq_t*)realloc(g->seq, (size_t)m * sizeof(asg_seq_t));
	}
	if ((uint32_t)sid >= g->n_seq) g->n_seq = sid + 1;
	g->seq[sid].len = len, g->seq[sid].del = !!del;
}

#define WITH_GRAPH(g, body) do { MabDev &d = mab_default_dev(); DGraph dg; mab_graph_upload(d, (g), dg); body; mab_graph_download(d, dg, (g)); dg_free(d, dg); d.sync(); } while (0)

void asg_arc_sort(asg_t *g) { WITH_GRAPH(g, { dg_arc_sort(d, dg); dg.is_srt = g->is_srt; }); }
void asg_arc_index(asg_t *g) { WITH_GRAPH(g, dg_arc_index(d, dg)); }
void asg_arc_rm(asg_t *g) { WITH_GRAPH(g, dg_arc_rm(d, dg, nullptr)); }
void asg_cleanup(asg_t *g) { WITH_GRAPH(g, dg_cleanup(d, dg)); }
void asg_symm(asg_t *g) { WITH_GRAPH(g, dg_symm(d, dg)); }
int asg_arc_del_multi(asg_t *g) { int r; WITH_GRAPH(g, r = (int)dg_del_multi(d, dg)); return r; }
int asg_arc_del_asymm(asg_t *g) { int r; WITH_GRAPH(g, r = (int)dg_del_asymm(d, dg)); return r; }
int asg_arc_del_trans(asg_t *g, int fuzz) { int r; WITH_GRAPH(g, r = (int)dg_del_trans(d, dg, (uint32_t)fuzz)); return r; }
int asg_arc_del_short(asg_t *g, float drop_ratio) { int r; WITH_GRAPH(g, r = (int)dg_del_short(d, dg, drop_ratio)); return r; }

void mab_set_verbose(int level) { mab_verbose = level; ma_verbose = level; }
void mab_count_del_trans_inner(int on) { mab_del_trans_count_inner = on; }

/* last asg_arc_del_trans kernel statistics, for tests and bench (drop-in level has no context object) */
void mab_last_del_trans(uint64_t *n_arc_in, uint64_t *inner, uint64_t *n_reduced, uint64_t *n_big, double *kernel_ms)
{
	if (n_arc_in) *n_arc_in = g_del_trans_stats.n_arc_in;
	if (inner) *inner = g_del_trans_stats.inner_iters;
	if (n_reduced) *n_reduced = g_del_trans_stats.n_reduced;
	if (n_big) *n_big = g_del_trans_stats.n_big;
	if (kernel_ms) *kernel_ms = g_del_trans_stats.kernel_ms;
}

}
