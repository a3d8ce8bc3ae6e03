// clean_dev.cu -- the order-dependent graph-cleaning passes of stage (iii) on the GPU, bit-exact with the
// reference's ascending-vertex sequential loops:
//   asg_cut_tip asg.c:238-254 | asg_cut_internal asg.c:256-272 | asg_cut_biloop asg.c:274-306 |
//   asg_pop_bubble asg.c:412-433 (asg_bub_pop1 asg.c:360-409, asg_bub_backtrack asg.c:338-357)
//
// Why this is not a plain parallel-for: iteration v reads deletion bits that iterations u < v may have
// written (asg_seq_del, asg.h:64-77), and the survey's probe shows a decide-on-snapshot variant changes
// the GFA.  The passes therefore run as SPECULATIVE PREFIX-COMMIT ROUNDS (DESIGN.md "sequential passes"):
//
//   state: everything below `lo` is final.  One round:
//   K1  every v >= lo evaluates its decision on the current state (read only).  A vertex that would act
//       (a "candidate") stamps every cell it reads or would write with atomicMin(tag[cell], v);
//       a cell is one read: its seq.del bit plus the del bits of both of its out-slabs.
//   K2  every v >= lo (candidate or not) walks its read set again (candidates: plus write set) and takes
//       the smallest stamp; if it is < v, an earlier candidate touches what v depends on, so v may change
//       once that candidate commits: v is "invalid".  x* = min invalid v (or infinity).
//   K3  candidates in [lo, x*) re-evaluate and apply their action.  They are pairwise cell-disjoint (every
//       candidate stamps all of its cells, so of two candidates sharing a cell the later one is invalid),
//       hence order-free, and nothing earlier can change them: exactly what the sequential loop does
//       for v < x*.  Non-candidates below x* are final no-ops.  lo = x*; repeat until no candidate is left.
//
// The smallest candidate is never invalid, so every round commits at least one action; on real graphs
// candidates are sparse and local and a handful of rounds finish a pass.
#include "clean_dev.cuh"
#include <cub/cub.cuh>

CleanStats g_clean_stats;

#define ET_MERGEABLE 0
#define ET_TIP       1
#define ET_MULTI_OUT 2
#define ET_MULTI_NEI 3

constexpr uint32_t NO_TAG = 0xffffffffu;

struct GV { // device view of the graph
	DArc *arc;
	const uint64_t *idx;
	uint32_t *seq;
	uint32_t n_vtx;
};

// visitors over cells
struct VisNone { __device__ __forceinline__ void operator()(uint32_t) const {} };
struct VisTag {
	uint32_t *tag, v;
	__device__ __forceinline__ void operator()(uint32_t cell) const { atomicMin(&tag[cell], v); }
};
struct VisMin {
	const uint32_t *tag; uint32_t m;
	__device__ __forceinline__ void operator()(uint32_t cell) { uint32_t t = tag[cell]; m = t < m ? t : m; }
};

// asg_is_utg_end (asg.c:204-222): looks at the live out-arcs of v^1 and of the single neighbour
template <class Vis>
__device__ __forceinline__ int is_utg_end(const GV &g, uint32_t v, uint64_t *lw, Vis &vis)
{
	const uint64_t iv = g.idx[v ^ 1];
	const uint32_t nv0 = (uint32_t)iv;
	const DArc *av = g.arc + (iv >> 32);
	uint32_t nv = 0, i0 = 0;
	vis(v >> 1);
	for (uint32_t i = 0; i < nv0; ++i)
		if (!(av[i].ol_del & MAB_DEL_BIT)) i0 = i, ++nv;
	if (nv == 0) return ET_TIP;
	if (nv > 1) return ET_MULTI_OUT;
	if (lw) *lw = av[i0].ul << 32 | av[i0].v;
	const uint32_t w = av[i0].v ^ 1;
	const uint64_t iw = g.idx[w];
	const uint32_t nw0 = (uint32_t)iw;
	const DArc *aw = g.arc + (iw >> 32);
	uint32_t nw = 0;
	vis(w >> 1);
	for (uint32_t i = 0; i < nw0; ++i)
		if (!(aw[i].ol_del & MAB_DEL_BIT)) ++nw;
	return nw != 1 ? ET_MULTI_NEI : ET_MERGEABLE;
}

// asg_extend (asg.c:224-236) without materialising the path: returns the end type, the last vertex pushed
// and the number of entries pushed (entry 0 is v itself).  `visit_chain(vertex)` sees every pushed vertex.
template <class Vis, class ChainFn>
__device__ __forceinline__ int extend(const GV &g, uint32_t v, int max_ext, Vis &vis, ChainFn chain, uint32_t *last)
{
	int ret;
	uint64_t lw = 0;
	chain(v);
	*last = v;
	do {
		ret = is_utg_end(g, v ^ 1, &lw, vis);
		if (ret != 0) break;
		v = (uint32_t)lw;
		chain(v);
		*last = v;
	} while (--max_ext > 0);
	return ret;
}

// cells written by asg_seq_del(read s) (asg.h:64-77): s itself and every read one of its arcs points to
template <class Vis>
__device__ __forceinline__ void seq_del_cells(const GV &g, uint32_t s, Vis &vis)
{
	vis(s);
	for (uint32_t k = 0; k < 2; ++k) {
		const uint64_t iv = g.idx[s << 1 | k];
		const DArc *av = g.arc + (iv >> 32);
		for (uint32_t i = 0; i < (uint32_t)iv; ++i) vis(av[i].v >> 1);
	}
}

__device__ __forceinline__ void arc_del(const GV &g, uint32_t v, uint32_t w, bool del) // asg_arc_del, asg.h:55-61
{
	const uint64_t iv = g.idx[v];
	DArc *av = g.arc + (iv >> 32);
	for (uint32_t i = 0; i < (uint32_t)iv; ++i)
		if (av[i].v == w) av[i].ol_del = del ? (av[i].ol_del | MAB_DEL_BIT) : (av[i].ol_del & ~MAB_DEL_BIT);
}

__device__ __forceinline__ void seq_del(const GV &g, uint32_t s) // asg_seq_del, asg.h:64-77
{
	g.seq[s] |= MAB_DEL_BIT;
	for (uint32_t k = 0; k < 2; ++k) {
		const uint32_t v = s << 1 | k;
		const uint64_t iv = g.idx[v];
		DArc *av = g.arc + (iv >> 32);
		for (uint32_t i = 0; i < (uint32_t)iv; ++i) {
			av[i].ol_del |= MAB_DEL_BIT;
			arc_del(g, av[i].v ^ 1, v ^ 1, true);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// The three short-unitig cutters share one skeleton; `Rule` supplies decision, cells and action.
//   eval<Vis>(g, v, vis)   -> true if v acts on the current state; vis sees every cell read
//   cells<Vis>(g, v, vis)  -> for an acting v: every cell its action writes
//   apply(g, v)            -> perform the action
// ---------------------------------------------------------------------------------------------
struct TipRule { // asg_cut_tip
	int max_ext;
	template <class Vis> __device__ bool eval(const GV &g, uint32_t v, Vis &vis) const
	{
		if (g.seq[v >> 1] & MAB_DEL_BIT) { vis(v >> 1); return false; }
		if (is_utg_end(g, v, nullptr, vis) != ET_TIP) return false;
		uint32_t last;
		return extend(g, v, max_ext, vis, [](uint32_t) {}, &last) != ET_MERGEABLE;
	}
	template <class Vis> __device__ void cells(const GV &g, uint32_t v, Vis &vis) const
	{
		VisNone none;
		uint32_t last;
		extend(g, v, max_ext, none, [&](uint32_t x) { seq_del_cells(g, x >> 1, vis); }, &last);
	}
	__device__ void apply(const GV &g, uint32_t v) const
	{ // the chain is fixed by the (stable) state before any deletion: collect first, delete after, like the reference
		uint32_t chain[64], n = 0, last;
		VisNone none;
		extend(g, v, max_ext < 63 ? max_ext : 63, none, [&](uint32_t x) { if (n < 64) chain[n++] = x; }, &last);
		for (uint32_t i = 0; i < n; ++i) seq_del(g, chain[i] >> 1);
	}
};

struct InternalRule { // asg_cut_internal
	int max_ext;
	template <class Vis> __device__ bool eval(const GV &g, uint32_t v, Vis &vis) const
	{
		if (g.seq[v >> 1] & MAB_DEL_BIT) { vis(v >> 1); return false; }
		if (is_utg_end(g, v, nullptr, vis) != ET_MULTI_NEI) return false;
		uint32_t last;
		return extend(g, v, max_ext, vis, [](uint32_t) {}, &last) == ET_MULTI_NEI;
	}
	template <class Vis> __device__ void cells(const GV &g, uint32_t v, Vis &vis) const
	{
		VisNone none;
		uint32_t last;
		extend(g, v, max_ext, none, [&](uint32_t x) { seq_del_cells(g, x >> 1, vis); }, &last);
	}
	__device__ void apply(const GV &g, uint32_t v) const
	{
		uint32_t chain[64], n = 0, last;
		VisNone none;
		extend(g, v, max_ext < 63 ? max_ext : 63, none, [&](uint32_t x) { if (n < 64) chain[n++] = x; }, &last);
		for (uint32_t i = 0; i < n; ++i) seq_del(g, chain[i] >> 1);
	}
};

struct BiloopRule { // asg_cut_biloop: v->...->x', w->v and w->x; drop w->x (and its complement) if it is the weaker one
	int max_ext;
	template <class Vis> __device__ bool find(const GV &g, uint32_t v, Vis &vis, uint32_t *w_out, uint32_t *x_out) const
	{
		if (g.seq[v >> 1] & MAB_DEL_BIT) { vis(v >> 1); return false; }
		if (is_utg_end(g, v, nullptr, vis) != ET_MULTI_NEI) return false;
		uint32_t last;
		if (extend(g, v, max_ext, vis, [](uint32_t) {}, &last) != ET_MULTI_OUT) return false;
		const uint32_t x = last ^ 1;
		uint32_t w = 0xffffffffu, ov = 0, ox = 0;
		{
			const uint64_t iv = g.idx[v ^ 1];
			const DArc *av = g.arc + (iv >> 32);
			for (uint32_t i = 0; i < (uint32_t)iv; ++i)
				if (!(av[i].ol_del & MAB_DEL_BIT)) w = av[i].v ^ 1;
		}
		if (w == 0xffffffffu) return false; // cannot happen: MULTI_NEI means exactly one live arc (asg.c:288 asserts it)
		vis(w >> 1);
		const uint64_t iw = g.idx[w];
		const DArc *aw = g.arc + (iw >> 32);
		for (uint32_t i = 0; i < (uint32_t)iw; ++i) {
			if (aw[i].ol_del & MAB_DEL_BIT) continue;
			if (aw[i].v == x) ox = aw[i].ol_del & ~MAB_DEL_BIT;
			if (aw[i].v == v) ov = aw[i].ol_del & ~MAB_DEL_BIT;
		}
		if (ov == 0 && ox == 0) return false;
		*w_out = w, *x_out = x;
		return ov > ox;
	}
	template <class Vis> __device__ bool eval(const GV &g, uint32_t v, Vis &vis) const
	{
		uint32_t w, x;
		return find(g, v, vis, &w, &x);
	}
	template <class Vis> __device__ void cells(const GV &g, uint32_t v, Vis &vis) const
	{
		VisNone none;
		uint32_t w, x;
		if (find(g, v, none, &w, &x)) vis(w >> 1), vis(x >> 1);
	}
	__device__ void apply(const GV &g, uint32_t v) const
	{
		VisNone none;
		uint32_t w, x;
		if (find(g, v, none, &w, &x)) arc_del(g, w, x, true), arc_del(g, x ^ 1, w ^ 1, true);
	}
};

template <class Rule>
__global__ void k_spec_eval(GV g, Rule rule, uint32_t lo, uint8_t *cand, uint32_t *tag, unsigned long long *n_cand)
{
	unsigned cnt = 0;
	for (uint32_t v = lo + blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x) {
		VisNone none;
		bool act = rule.eval(g, v, none);
		cand[v] = act;
		if (act) {
			VisTag t{tag, v};
			rule.eval(g, v, t);
			rule.cells(g, v, t);
			++cnt;
		}
	}
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_cand, (unsigned long long)cnt);
}

template <class Rule>
__global__ void k_spec_check(GV g, Rule rule, uint32_t lo, const uint8_t *cand, const uint32_t *tag, unsigned long long *xstar)
{
	uint32_t bad = NO_TAG;
	for (uint32_t v = lo + blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x) {
		VisMin m{tag, NO_TAG};
		rule.eval(g, v, m);
		if (cand[v]) rule.cells(g, v, m);
		if (m.m < v && v < bad) bad = v;
	}
	bad = __reduce_min_sync(0xffffffffu, bad);
	if ((threadIdx.x & 31) == 0 && bad != NO_TAG) atomicMin(xstar, (unsigned long long)bad);
}

template <class Rule>
__global__ void k_spec_commit(GV g, Rule rule, uint32_t lo, uint32_t hi, const uint8_t *cand, unsigned long long *n_done)
{
	unsigned cnt = 0;
	for (uint32_t v = lo + blockIdx.x * blockDim.x + threadIdx.x; v < hi; v += gridDim.x * blockDim.x)
		if (cand[v]) { rule.apply(g, v); ++cnt; }
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_done, (unsigned long long)cnt);
}

// Windowed rounds (experimental, MAB_SPEC_WINDOW=1): a round only looks at the vertices [lo, lo + W).  Vertices beyond
// the window neither stamp nor commit in that round, and a stamp only ever invalidates LARGER vertices, so the prefix
// that commits is still exactly the sequential outcome; what changes is the cost of a round, O(W) instead of
// O(n_vtx - lo).  W shrinks towards the distance the last round advanced and grows back when a whole window commits:
// dense conflict chains (bubbles along a contig with sorted ids advance a few vertices per round) stop paying for
// a full-graph evaluation each time.  The kernels are unchanged: they loop to GV::n_vtx, which is set to the window end.
struct SpecWindow {
	bool on; uint32_t n_vtx, W;
	explicit SpecWindow(uint32_t n) : n_vtx(n), W(n) { const char *e = getenv("MAB_SPEC_WINDOW"); on = e && atoi(e) != 0; }
	uint32_t end(uint32_t lo) const { return !on || (uint64_t)lo + W >= n_vtx ? n_vtx : lo + W; }
	void advanced(uint32_t lo, uint32_t hi, uint32_t end) {
		if (!on) return;
		if (hi >= end) W = (uint64_t)W * 4 >= n_vtx ? n_vtx : W * 4;                 // the whole window went through
		else { const uint64_t w = 16ull * (hi - lo + 1); W = w < 4096 ? 4096 : (w < W ? (uint32_t)w : W); }
	}
};

template <class Rule>
static uint32_t run_spec_rounds(MabDev &d, DGraph &g, Rule rule)
{
	const uint32_t n_vtx = g.n_seq * 2;
	uint32_t lo = 0, total = 0, rounds = 0;
	if (n_vtx == 0 || g.n_arc == 0) { g_clean_stats.rounds = 0, g_clean_stats.committed = 0; return 0; }
	GV gv{g.arc, g.idx, g.seq, n_vtx};
	uint8_t *cand = mab_alloc<uint8_t>(d, n_vtx);
	uint32_t *tag = mab_alloc<uint32_t>(d, g.n_seq);
	MAB_CUDA(cudaMemsetAsync(tag, 0xff, (size_t)g.n_seq * 4, d.stream));
	SpecWindow win(n_vtx);
	while (lo < n_vtx) {
		const uint32_t end = win.end(lo);
		GV gw = gv;
		gw.n_vtx = end;                                              // the eval / check kernels stop at the window end
		const unsigned grid = mab_grid(end - lo, 128);
		d.zero_scal(SC_COUNT, 2);
		MAB_CUDA(cudaMemsetAsync(d.d_scal + SC_MIN, 0xff, 8, d.stream));
		MAB_LAUNCH(d, k_spec_eval<Rule>, grid, 128, 0, gw, rule, lo, cand, tag, d.d_scal + SC_COUNT);
		if (d.get_scal(SC_COUNT) == 0) {                             // nobody in [lo, end) would act (and nobody stamped)
			if (end == n_vtx) break;
			win.advanced(lo, end, end);
			lo = end;
			continue;
		}
		MAB_LAUNCH(d, k_spec_check<Rule>, grid, 128, 0, gw, rule, lo, cand, tag, d.d_scal + SC_MIN);
		unsigned long long xs = d.get_scal(SC_MIN);
		uint32_t hi = xs >= end ? end : (uint32_t)xs;
		MAB_LAUNCH(d, k_spec_commit<Rule>, mab_grid(hi - lo, 128), 128, 0, gv, rule, lo, hi, cand, d.d_scal + SC_NSEL);
		MAB_CUDA(cudaMemsetAsync(tag, 0xff, (size_t)g.n_seq * 4, d.stream));
		total += (uint32_t)d.get_scal(SC_NSEL);
		win.advanced(lo, hi, end);
		lo = hi;
		++rounds;
	}
	d.free(cand); d.free(tag);
	g_clean_stats.rounds = rounds, g_clean_stats.committed = total;
	return total;
}

uint32_t dg_cut_tip(MabDev &d, DGraph &g, int max_ext)
{
	uint32_t cnt = run_spec_rounds(d, g, TipRule{max_ext});
	if (cnt > 0) dg_cleanup(d, g);
	if (mab_verbose >= 1) fprintf(stderr, "[M::%s] cut %d tips\n", "asg_cut_tip", cnt);
	return cnt;
}

uint32_t dg_cut_internal(MabDev &d, DGraph &g, int max_ext)
{
	uint32_t cnt = run_spec_rounds(d, g, InternalRule{max_ext});
	if (cnt > 0) dg_cleanup(d, g);
	if (mab_verbose >= 1) fprintf(stderr, "[M::%s] cut %d internal sequences\n", "asg_cut_internal", cnt);
	return cnt;
}

uint32_t dg_cut_biloop(MabDev &d, DGraph &g, int max_ext)
{
	uint32_t cnt = run_spec_rounds(d, g, BiloopRule{max_ext});
	if (cnt > 0) dg_cleanup(d, g);
	if (mab_verbose >= 1) fprintf(stderr, "[M::%s] cut %d small bi-loops\n", "asg_cut_biloop", cnt);
	return cnt;
}

// ---------------------------------------------------------------------------------------------
// Bubble popping (asg.c:312-433).  asg_bub_pop1 is a bounded Kahn-style traversal from a source v0 with
// >= 2 live out-arcs; per visited vertex it keeps {best parent p, distance d, read count c, pending in-arcs r}.
// The reference indexes one n_vtx-sized array by vertex and resets the touched entries afterwards; here each
// traversal owns a small open-addressing table vertex -> {p,d,c,r} in a per-thread scratch slot, which is the
// same map restricted to the visited set.  One thread walks one source (the walk is a LIFO-ordered pointer
// chase); sources are independent within a round of the prefix-commit scheme above.  Cells of a traversal:
// the source's read and the read of every visited vertex (all state read or written lives there).
// ---------------------------------------------------------------------------------------------
struct BubSlots {
	uint32_t *hkey, *hp, *hd, *hc, *hr;  // [n_slot][hcap]
	uint32_t *b, *bslot, *S;             // [n_slot][bcap]
	uint32_t *e;                         // [n_slot][ecap]
	uint32_t bcap, ecap, hcap, n_slot;
};

struct BubWalk { uint32_t nb, ne, nT, sink; };

constexpr uint32_t BUB_EMPTY = 0xffffffffu;

__device__ __forceinline__ uint32_t bub_find(const uint32_t *hkey, uint32_t hmask, uint32_t key)
{
	uint32_t h = (key * 2654435761u) >> 9 & hmask;
	while (hkey[h] != key) h = (h + 1) & hmask; // the key is known to be present
	return h;
}

// returns 1 = bubble resolved (backtrack applies), 0 = nothing to pop, -1 = scratch too small
__device__ int bub_walk(const GV &g, uint32_t v0, uint32_t max_dist, const BubSlots &sl, uint32_t slot, BubWalk *out)
{
	uint32_t *hkey = sl.hkey + (size_t)slot * sl.hcap, *hp = sl.hp + (size_t)slot * sl.hcap, *hd = sl.hd + (size_t)slot * sl.hcap;
	uint32_t *hc = sl.hc + (size_t)slot * sl.hcap, *hr = sl.hr + (size_t)slot * sl.hcap;
	uint32_t *b = sl.b + (size_t)slot * sl.bcap, *bslot = sl.bslot + (size_t)slot * sl.bcap, *S = sl.S + (size_t)slot * sl.bcap;
	uint32_t *e = sl.e + (size_t)slot * sl.ecap;
	const uint32_t hmask = sl.hcap - 1;
	uint32_t nb = 0, ne = 0, nT = 0, nS = 0, n_pending = 0;
	int ret = 0;
	S[nS++] = v0;
	do {
		const uint32_t v = S[--nS];
		uint32_t d = 0, c = 0;
		if (v != v0) { uint32_t h = bub_find(hkey, hmask, v); d = hd[h], c = hc[h]; }
		const uint64_t iv = g.idx[v];
		const uint32_t nv = (uint32_t)iv, off = (uint32_t)(iv >> 32);
		const DArc *av = g.arc + off;
		uint32_t i;
		for (i = 0; i < nv; ++i) {
			const uint32_t w = av[i].v, l = (uint32_t)av[i].ul;
			if (w == v0) goto done;                        // a cycle through the source (tested before the del bit, asg.c:377)
			if (av[i].ol_del & MAB_DEL_BIT) continue;
			if (ne == sl.ecap) { ret = -1; goto done; }
			e[ne++] = off + i;
			if (d + l > max_dist) break;                   // too far
			uint32_t h = (w * 2654435761u) >> 9 & hmask;
			while (hkey[h] != BUB_EMPTY && hkey[h] != w) h = (h + 1) & hmask;
			if (hkey[h] == BUB_EMPTY) {                    // first visit
				if (nb == sl.bcap) { ret = -1; goto done; }
				hkey[h] = w; bslot[nb] = h; b[nb++] = w;
				hp[h] = v, hd[h] = d + l, hc[h] = 0;
				uint32_t r = 0;                            // count_out(w^1): live arcs only
				const uint64_t ix = g.idx[w ^ 1];
				const DArc *ax = g.arc + (ix >> 32);
				for (uint32_t k = 0; k < (uint32_t)ix; ++k) r += !(ax[k].ol_del & MAB_DEL_BIT);
				hr[h] = r;
				++n_pending;
			} else {
				if (c + 1 > hc[h] || (c + 1 == hc[h] && d + l > hd[h])) hp[h] = v;
				if (c + 1 > hc[h]) hc[h] = c + 1;
				if (d + l < hd[h]) hd[h] = d + l;
			}
			hr[h] = (hr[h] - 1) & 0x7fffffffu;
			if (hr[h] == 0) {
				if ((uint32_t)g.idx[w]) S[nS++] = w;       // nS <= nb + 1 <= bcap: every vertex is pushed at most once
				else ++nT;                                 // a tip
				--n_pending;
			}
		}
		if (i < nv || nS == 0) goto done;
	} while (nS > 1 || n_pending);
	ret = 1;
	out->sink = S[0];
done:
	out->nb = nb, out->ne = ne, out->nT = nT;
	return ret;
}

__device__ __forceinline__ void bub_reset(const BubSlots &sl, uint32_t slot, uint32_t nb)
{
	uint32_t *hkey = sl.hkey + (size_t)slot * sl.hcap;
	const uint32_t *bslot = sl.bslot + (size_t)slot * sl.bcap;
	for (uint32_t i = 0; i < nb; ++i) hkey[bslot[i]] = BUB_EMPTY;
}

// asg_bub_backtrack (asg.c:338-357)
__device__ void bub_backtrack(const GV &g, uint32_t v0, const BubSlots &sl, uint32_t slot, const BubWalk &w)
{
	const uint32_t *hkey = sl.hkey + (size_t)slot * sl.hcap, *hp = sl.hp + (size_t)slot * sl.hcap;
	const uint32_t *b = sl.b + (size_t)slot * sl.bcap, *e = sl.e + (size_t)slot * sl.ecap;
	for (uint32_t i = 0; i < w.nb; ++i) g.seq[b[i] >> 1] |= MAB_DEL_BIT;
	for (uint32_t i = 0; i < w.ne; ++i) {
		DArc *a = g.arc + e[i];
		a->ol_del |= MAB_DEL_BIT;
		arc_del(g, a->v ^ 1, (uint32_t)(a->ul >> 32) ^ 1, true);
	}
	uint32_t v = w.sink;
	do {
		const uint32_t u = hp[bub_find(hkey, sl.hcap - 1, v)];
		g.seq[v >> 1] &= ~MAB_DEL_BIT;
		arc_del(g, u, v, false);
		arc_del(g, v ^ 1, u ^ 1, false);
		v = u;
	} while (v != v0);
}

// the outer-loop test of asg_pop_bubble (asg.c:420-426) plus the guards of asg_bub_pop1 (asg.c:364-365)
__device__ __forceinline__ bool bub_is_source(const GV &g, uint32_t v)
{
	const uint64_t iv = g.idx[v];
	const uint32_t nv = (uint32_t)iv;
	if (nv < 2 || (g.seq[v >> 1] & MAB_DEL_BIT)) return false;
	const DArc *av = g.arc + (iv >> 32);
	uint32_t live = 0;
	for (uint32_t i = 0; i < nv; ++i) live += !(av[i].ol_del & MAB_DEL_BIT);
	return live > 1;
}

__global__ void k_bub_sources(GV g, uint32_t lo, uint32_t *src, unsigned long long *n_src)
{
	for (uint32_t v = lo + blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x)
		if (bub_is_source(g, v)) src[atomicAdd(n_src, 1ull)] = v;
}

__global__ void k_bub_eval(GV g, uint32_t max_dist, BubSlots sl, const uint32_t *src, uint32_t n_src, uint8_t *cand, uint32_t *tag,
                           unsigned long long *n_cand, unsigned long long *overflow)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= sl.n_slot) return;
	for (uint32_t k = slot; k < n_src; k += sl.n_slot) {
		const uint32_t v0 = src[k];
		BubWalk w;
		int r = bub_walk(g, v0, max_dist, sl, slot, &w);
		cand[v0] = r == 1;
		if (r == 1) {
			const uint32_t *b = sl.b + (size_t)slot * sl.bcap;
			atomicMin(&tag[v0 >> 1], v0);
			for (uint32_t i = 0; i < w.nb; ++i) atomicMin(&tag[b[i] >> 1], v0);
			atomicAdd(n_cand, 1ull);
		} else if (r < 0) atomicAdd(overflow, 1ull);
		bub_reset(sl, slot, w.nb);
	}
}

// non-sources read only their own read's state
__global__ void k_bub_check_own(GV g, uint32_t lo, const uint32_t *tag, unsigned long long *xstar)
{
	uint32_t bad = NO_TAG;
	for (uint32_t v = lo + blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x)
		if (tag[v >> 1] < v && v < bad) bad = v;
	bad = __reduce_min_sync(0xffffffffu, bad);
	if ((threadIdx.x & 31) == 0 && bad != NO_TAG) atomicMin(xstar, (unsigned long long)bad);
}

__global__ void k_bub_check_walk(GV g, uint32_t max_dist, BubSlots sl, const uint32_t *src, uint32_t n_src, const uint32_t *tag, unsigned long long *xstar)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= sl.n_slot) return;
	for (uint32_t k = slot; k < n_src; k += sl.n_slot) {
		const uint32_t v0 = src[k];
		BubWalk w;
		bub_walk(g, v0, max_dist, sl, slot, &w);
		const uint32_t *b = sl.b + (size_t)slot * sl.bcap;
		uint32_t m = tag[v0 >> 1];
		for (uint32_t i = 0; i < w.nb; ++i) { uint32_t t = tag[b[i] >> 1]; m = t < m ? t : m; }
		if (m < v0) atomicMin(xstar, (unsigned long long)v0);
		bub_reset(sl, slot, w.nb);
	}
}

__global__ void k_bub_commit(GV g, uint32_t max_dist, BubSlots sl, const uint32_t *src, uint32_t n_src, const uint8_t *cand, uint32_t hi,
                             unsigned long long *n_pop, unsigned long long *n_tip)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= sl.n_slot) return;
	for (uint32_t k = slot; k < n_src; k += sl.n_slot) {
		const uint32_t v0 = src[k];
		if (v0 >= hi || !cand[v0]) continue;
		BubWalk w;
		if (bub_walk(g, v0, max_dist, sl, slot, &w) == 1) {
			bub_backtrack(g, v0, sl, slot, w);
			atomicAdd(n_pop, 1ull);
			if (w.nT) atomicAdd(n_tip, (unsigned long long)w.nT);
		}
		bub_reset(sl, slot, w.nb);
	}
}

// ---- "excuse" variant of the validity check (experimental, MAB_BUB_EXCUSE=1) ----------------------------------------
// With the plain prefix rule every pop costs a round on genome-ordered ids: the pop at source A stamps the reads of its
// region, the complement-strand twin of the bubble (source = sink(A)^1, a slightly larger id) finds those stamps and is
// invalid, so x* lands right behind A.  But that twin -- like every other vertex on a read of A's walk set except A^1
// and sink(A) -- has all of its live out-arcs (on the complement strand: the complements of all its live in-arcs)
// inside A's region: once A has popped, at most the restored path arc is left, it is no longer a source, and in this
// deletion-only pass (no multi-arcs, the graph is symmetric) it never becomes one again.  Such a vertex is a
// guaranteed no-op and need not stop the prefix.  Non-sources are final no-ops for the same reason and are not
// checked at all.  oracle/spec_sim.c is the CPU model of exactly this rule: identical final state on every parity
// set, 3713 -> 174 rounds on a 300 K-read bubble-dense set.
__global__ void k_bub_check_walk2(GV g, uint32_t max_dist, BubSlots sl, const uint32_t *src, uint32_t n_src, const uint32_t *tag,
                                  uint32_t *mn, uint32_t *sink)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= sl.n_slot) return;
	for (uint32_t k = slot; k < n_src; k += sl.n_slot) {
		const uint32_t v0 = src[k];
		BubWalk w;
		const int r = bub_walk(g, v0, max_dist, sl, slot, &w);
		const uint32_t *b = sl.b + (size_t)slot * sl.bcap;
		uint32_t m = tag[v0 >> 1];
		for (uint32_t i = 0; i < w.nb; ++i) { uint32_t t = tag[b[i] >> 1]; m = t < m ? t : m; }
		mn[v0] = m;                                  // smallest stamp on the walk set (valid iff >= v0)
		sink[v0] = r == 1 ? w.sink : NO_TAG;
		bub_reset(sl, slot, w.nb);
	}
}

__global__ void k_bub_xstar2(const uint32_t *src, uint32_t n_src, const uint32_t *tag, const uint8_t *cand, const uint32_t *mn, const uint32_t *sink,
                             unsigned long long *xstar)
{
	uint32_t bad = NO_TAG;
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_src; k += gridDim.x * blockDim.x) {
		const uint32_t v = src[k];
		if (mn[v] >= v) continue;                    // valid
		const uint32_t A = tag[v >> 1];              // smallest candidate that stamped v's own read (it is a source of this round)
		const bool excused = A < v && cand[A] && mn[A] >= A && v != (A ^ 1) && v != sink[A];
		if (!excused && v < bad) bad = v;
	}
	bad = __reduce_min_sync(0xffffffffu, bad);
	if ((threadIdx.x & 31) == 0 && bad != NO_TAG) atomicMin(xstar, (unsigned long long)bad);
}

// commit of the excuse variant: below x* there may be invalid (excused) candidates, which must not act
__global__ void k_bub_commit2(GV g, uint32_t max_dist, BubSlots sl, const uint32_t *src, uint32_t n_src, const uint8_t *cand, const uint32_t *mn, uint32_t hi,
                              unsigned long long *n_pop, unsigned long long *n_tip)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= sl.n_slot) return;
	for (uint32_t k = slot; k < n_src; k += sl.n_slot) {
		const uint32_t v0 = src[k];
		if (v0 >= hi || !cand[v0] || mn[v0] < v0) continue;
		BubWalk w;
		if (bub_walk(g, v0, max_dist, sl, slot, &w) == 1) {
			bub_backtrack(g, v0, sl, slot, w);
			atomicAdd(n_pop, 1ull);
			if (w.nT) atomicAdd(n_tip, (unsigned long long)w.nT);
		}
		bub_reset(sl, slot, w.nb);
	}
}

static void bub_slots_alloc(MabDev &d, BubSlots &sl, uint32_t n_slot, uint32_t bcap)
{
	sl.n_slot = n_slot, sl.bcap = bcap, sl.ecap = bcap * 4;
	sl.hcap = 1; while (sl.hcap < 2 * bcap) sl.hcap <<= 1;
	size_t nh = (size_t)n_slot * sl.hcap, nb = (size_t)n_slot * sl.bcap, ne = (size_t)n_slot * sl.ecap;
	sl.hkey = mab_alloc<uint32_t>(d, nh); sl.hp = mab_alloc<uint32_t>(d, nh); sl.hd = mab_alloc<uint32_t>(d, nh);
	sl.hc = mab_alloc<uint32_t>(d, nh); sl.hr = mab_alloc<uint32_t>(d, nh);
	sl.b = mab_alloc<uint32_t>(d, nb); sl.bslot = mab_alloc<uint32_t>(d, nb); sl.S = mab_alloc<uint32_t>(d, nb);
	sl.e = mab_alloc<uint32_t>(d, ne);
	MAB_CUDA(cudaMemsetAsync(sl.hkey, 0xff, nh * 4, d.stream));
}

static void bub_slots_free(MabDev &d, BubSlots &sl)
{
	d.free(sl.hkey); d.free(sl.hp); d.free(sl.hd); d.free(sl.hc); d.free(sl.hr);
	d.free(sl.b); d.free(sl.bslot); d.free(sl.S); d.free(sl.e);
}

uint64_t dg_pop_bubble(MabDev &d, DGraph &g, int max_dist)
{
	const uint32_t n_vtx = g.n_seq * 2;
	uint64_t n_pop = 0, n_tip = 0;
	uint32_t rounds = 0;
	if (!g.is_symm) dg_symm(d, g);
	if (n_vtx && g.n_arc) {
		GV gv{g.arc, g.idx, g.seq, n_vtx};
		uint8_t *cand = mab_alloc<uint8_t>(d, n_vtx);
		uint32_t *tag = mab_alloc<uint32_t>(d, g.n_seq);
		uint32_t *src = mab_alloc<uint32_t>(d, n_vtx);
		MAB_CUDA(cudaMemsetAsync(tag, 0xff, (size_t)g.n_seq * 4, d.stream));
		MAB_CUDA(cudaMemsetAsync(cand, 0, n_vtx, d.stream));
		static const bool excuse = getenv("MAB_BUB_EXCUSE") && atoi(getenv("MAB_BUB_EXCUSE")) != 0;
		uint32_t *mn = excuse ? mab_alloc<uint32_t>(d, n_vtx) : nullptr, *sink = excuse ? mab_alloc<uint32_t>(d, n_vtx) : nullptr;
		BubSlots sl;
		uint32_t bcap = 256, n_slot = 4096;
		bub_slots_alloc(d, sl, n_slot, bcap);
		uint32_t lo = 0;
		SpecWindow win(n_vtx);
		while (lo < n_vtx) {
			const uint32_t end = win.end(lo);
			GV gw = gv;
			gw.n_vtx = end;                                          // sources and own-cell checks of [lo, end) only
			d.zero_scal(SC_COUNT, 4); // COUNT (candidates), NSEL, BIG (overflow), AUX (sources)
			MAB_CUDA(cudaMemsetAsync(d.d_scal + SC_MIN, 0xff, 8, d.stream));
			MAB_LAUNCH(d, k_bub_sources, mab_grid(end - lo, 256), 256, 0, gw, lo, src, d.d_scal + SC_AUX);
			uint32_t n_src = (uint32_t)d.get_scal(SC_AUX);
			if (n_src == 0) {
				if (end == n_vtx) break;
				win.advanced(lo, end, end);
				lo = end;
				continue;
			}
			const unsigned wgrid = (sl.n_slot + 63) / 64;
			MAB_LAUNCH(d, k_bub_eval, wgrid, 64, 0, gv, (uint32_t)max_dist, sl, src, n_src, cand, tag, d.d_scal + SC_COUNT, d.d_scal + SC_BIG);
			uint64_t n_cand = d.get_scal(SC_COUNT);
			if (d.h_scal[SC_BIG]) { // a traversal outgrew its scratch slot: enlarge and redo the round
				bub_slots_free(d, sl);
				bcap *= 4;
				if (n_slot > 64) n_slot /= 4;
				if ((uint64_t)bcap > (uint64_t)n_vtx * 4) { fprintf(stderr, "[E::miniasm_b200] bubble scratch overflow\n"); exit(75); }
				bub_slots_alloc(d, sl, n_slot, bcap);
				MAB_CUDA(cudaMemsetAsync(tag, 0xff, (size_t)g.n_seq * 4, d.stream));
				continue;
			}
			if (n_cand == 0) {                                       // no source of the window pops anything (nobody stamped)
				if (end == n_vtx) break;
				win.advanced(lo, end, end);
				lo = end;
				continue;
			}
			if (excuse) { // non-sources unchecked, excusable invalid sources do not stop the prefix (see k_bub_check_walk2)
				MAB_LAUNCH(d, k_bub_check_walk2, wgrid, 64, 0, gv, (uint32_t)max_dist, sl, src, n_src, tag, mn, sink);
				MAB_LAUNCH(d, k_bub_xstar2, mab_grid(n_src, 256), 256, 0, src, n_src, tag, cand, mn, sink, d.d_scal + SC_MIN);
			} else {
				MAB_LAUNCH(d, k_bub_check_own, mab_grid(end - lo, 256), 256, 0, gw, lo, tag, d.d_scal + SC_MIN);
				MAB_LAUNCH(d, k_bub_check_walk, wgrid, 64, 0, gv, (uint32_t)max_dist, sl, src, n_src, tag, d.d_scal + SC_MIN);
			}
			unsigned long long xs = d.get_scal(SC_MIN);
			uint32_t hi = xs >= end ? end : (uint32_t)xs;
			d.zero_scal(SC_TMP0, 2);
			if (excuse) MAB_LAUNCH(d, k_bub_commit2, wgrid, 64, 0, gv, (uint32_t)max_dist, sl, src, n_src, cand, mn, hi, d.d_scal + SC_TMP0, d.d_scal + SC_TMP0 + 1);
			else MAB_LAUNCH(d, k_bub_commit, wgrid, 64, 0, gv, (uint32_t)max_dist, sl, src, n_src, cand, hi, d.d_scal + SC_TMP0, d.d_scal + SC_TMP0 + 1);
			MAB_CUDA(cudaMemsetAsync(tag, 0xff, (size_t)g.n_seq * 4, d.stream));
			n_pop += d.get_scal(SC_TMP0);
			n_tip += d.h_scal[SC_TMP0 + 1];
			win.advanced(lo, hi, end);
			lo = hi;
			++rounds;
		}
		bub_slots_free(d, sl);
		d.free(cand); d.free(tag); d.free(src); d.free(mn); d.free(sink);
	}
	g_clean_stats.rounds = rounds, g_clean_stats.committed = (uint32_t)n_pop;
	if (n_pop) dg_cleanup(d, g);
	if (mab_verbose >= 1) fprintf(stderr, "[M::%s] popped %d bubbles and trimmed %d tips\n", "asg_pop_bubble", (uint32_t)n_pop, (uint32_t)n_tip);
	return (n_pop & 0xffffffffull) | n_tip << 32;
}

// ---------------------------------------------------------------------------------------------
// ma_ug_gen (asm.c:121-210): unitigs = maximal chains of arcs w->x with out(w) == 1 and out(x^1) == 1.
//
// Sequential reference: scan v ascending; the first unvisited v with arcs seeds a unitig, walks forward to
// the chain end, then backward to the chain start, and marks both strands of everything on it.  Consequences
// used here: (1) every oriented vertex lies on exactly one chain of the successor function F (and the
// complement strand on the mirrored chain); (2) a chain pair {C, rc(C)} yields ONE unitig, oriented like the
// chain that holds the smallest seedable vertex (not deleted, has arcs) of C u rc(C), and unitigs are numbered
// by ascending seed; (3) a chain that closes on itself is circular and listed from its seed.
//
// GPU shape: F/B successor arrays -> pointer doubling along both directions (min seedable id, hop count and
// arc-length prefix sums to the chain head) -> seeds flagged, exclusive scans give unitig numbers and item
// offsets -> every vertex of an emitted chain writes its own item.  O(n log n) work, log n launches.
// The doubling needs F and B to be mutually inverse, which holds on a symmetric graph (every pass after
// transitive reduction leaves one); otherwise (e.g. `-S5 -p ug` on the raw graph) one thread replays the
// reference's walk literally.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t NONE = 0xffffffffu;

struct UgArrays {
	uint32_t *F, *B;          // successor / predecessor on the chain (NONE at the ends)
	uint32_t *jf, *jb;        // doubling pointers
	uint32_t *mf, *mb;        // min seedable vertex over the covered span, forward / backward
	uint32_t *rk;             // hops to the head (paths) or to the seed (cycles)
	uint64_t *ps;             // sum of arc lengths from the head/seed up to (excluding) this vertex
	uint32_t *cnt;            // hops to the tail, forward
	uint32_t *seedmin;        // min seedable vertex of the vertex's own chain
};

__device__ __forceinline__ uint32_t ug_fwd(const GV &g, uint32_t w)
{
	const uint64_t iw = g.idx[w];
	if ((uint32_t)iw != 1) return NONE;
	const uint32_t x = g.arc[iw >> 32].v;
	return (uint32_t)g.idx[x ^ 1] == 1 ? x : NONE;
}
__device__ __forceinline__ uint32_t ug_bwd(const GV &g, uint32_t x)
{
	const uint64_t ix = g.idx[x ^ 1];
	if ((uint32_t)ix != 1) return NONE;
	const uint32_t w = g.arc[ix >> 32].v ^ 1;
	return (uint32_t)g.idx[w] == 1 ? w : NONE;
}

__global__ void k_ug_links(GV g, uint32_t *F, uint32_t *B)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x)
		F[v] = ug_fwd(g, v), B[v] = ug_bwd(g, v);
}

__global__ void k_ug_consistent(uint32_t n_vtx, const uint32_t *F, const uint32_t *B, unsigned long long *bad)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vtx; v += gridDim.x * blockDim.x) {
		bool ok = true;
		if (F[v] != NONE && B[F[v]] != v) ok = false;
		if (B[v] != NONE && F[B[v]] != v) ok = false;
		if (F[v] != NONE && F[v] == (v ^ 1)) ok = false; // a read chained to its own complement: let the literal walk handle it
		if (!ok) atomicAdd(bad, 1ull);
	}
}

__global__ void k_ug_init(GV g, UgArrays a)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x) {
		const bool seedable = !(g.seq[v >> 1] & MAB_DEL_BIT) && (uint32_t)g.idx[v] != 0;
		const uint32_t m = seedable ? v : NONE;
		a.jf[v] = a.F[v], a.jb[v] = a.B[v];
		a.mf[v] = m, a.mb[v] = m;
	}
}

// one doubling step for the minima (double-buffered by the caller through jf/jb + mf/mb copies)
__global__ void k_ug_min_step(uint32_t n_vtx, const uint32_t *jf, const uint32_t *jb, const uint32_t *mf, const uint32_t *mb,
                              uint32_t *jf2, uint32_t *jb2, uint32_t *mf2, uint32_t *mb2)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vtx; v += gridDim.x * blockDim.x) {
		uint32_t f = jf[v], b = jb[v], x = mf[v], y = mb[v];
		if (f != NONE) { uint32_t t = mf[f]; x = t < x ? t : x; f = jf[f]; }
		if (b != NONE) { uint32_t t = mb[b]; y = t < y ? t : y; b = jb[b]; }
		jf2[v] = f, jb2[v] = b, mf2[v] = x, mb2[v] = y;
	}
}

// after the min doubling: seedmin = min(forward, backward); cycles are cut open just before their seed
__global__ void k_ug_cut(GV g, UgArrays a, const uint32_t *jb_final)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x) {
		const uint32_t m = a.mf[v] < a.mb[v] ? a.mf[v] : a.mb[v];
		a.seedmin[v] = m;
		// a vertex still holding a live backward pointer after ceil(log2 n)+1 doublings sits on a cycle
		const bool cyc = jb_final[v] != NONE;
		uint32_t b = a.B[v];
		if (cyc && v == m) b = NONE;               // the seed becomes the head of its (opened) cycle
		a.jb[v] = b;
		a.rk[v] = b == NONE ? 0 : 1;
		uint64_t l = 0;
		if (b != NONE) l = (uint32_t)g.arc[g.idx[b] >> 32].ul; // length of the arc b -> v (b has exactly one arc)
		a.ps[v] = l;
		uint32_t f = a.F[v];
		if (cyc && f == m) f = NONE;               // ... and its predecessor the tail
		a.jf[v] = f;
		a.cnt[v] = f == NONE ? 0 : 1;
	}
}

__global__ void k_ug_rank_step(uint32_t n_vtx, const uint32_t *jb, const uint32_t *rk, const uint64_t *ps, const uint32_t *jf, const uint32_t *cnt,
                               uint32_t *jb2, uint32_t *rk2, uint64_t *ps2, uint32_t *jf2, uint32_t *cnt2)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vtx; v += gridDim.x * blockDim.x) {
		uint32_t b = jb[v], r = rk[v], f = jf[v], c = cnt[v];
		uint64_t p = ps[v];
		if (b != NONE) r += rk[b], p += ps[b], b = jb[b];
		if (f != NONE) c += cnt[f], f = jf[f];
		jb2[v] = b, rk2[v] = r, ps2[v] = p, jf2[v] = f, cnt2[v] = c;
	}
}

// head of an emitted chain: rank 0 and its chain wins over the mirrored chain (or is its own mirror)
__global__ void k_ug_heads(GV g, UgArrays a, const uint32_t *is_cyc, uint32_t *flag_seed, uint32_t *n_items_at_seed)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x) {
		// v is "the seed" iff it is the smallest seedable vertex of its chain pair and lies on the chain
		const uint32_t m = a.seedmin[v], mr = a.seedmin[v ^ 1];
		const bool seed = m != NONE && v == m && m < mr; // m == mr impossible: the two chains are disjoint vertex sets
		flag_seed[v] = seed;
		n_items_at_seed[v] = seed ? a.rk[v] + a.cnt[v] + 1 : 0; // chain length = hops to head + hops to tail + 1
	}
}

__global__ void k_ug_emit(GV g, UgArrays a, const uint32_t *is_cyc, const uint32_t *utg_of_seed, const uint32_t *first_of_seed,
                          uint64_t *items, DUtgMeta *meta)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vtx; v += gridDim.x * blockDim.x) {
		const uint32_t m = a.seedmin[v];
		if (m == NONE || !(m < a.seedmin[v ^ 1])) continue; // chain not emitted in this orientation
		const uint32_t u = utg_of_seed[m];
		const bool cyc = is_cyc[v] != 0;
		const bool tail = a.cnt[v] == 0;
		uint32_t l;
		if (!tail || cyc) l = (uint32_t)g.arc[g.idx[v] >> 32].ul; // arc to the successor (for a cycle: also tail -> seed)
		else l = g.seq[v >> 1] & 0x7fffffffu;                       // the last read contributes its full length
		items[first_of_seed[m] + a.rk[v]] = (uint64_t)v << 32 | l;
		if (tail) { // the tail knows the total: prefix up to itself + its own item
			DUtgMeta mt;
			mt.len = (uint32_t)((a.ps[v] + l) & 0x7fffffffu);
			mt.circ = cyc;
			mt.n = a.rk[v] + 1;
			mt.first = first_of_seed[m];
			mt.end = cyc ? NONE : (v ^ 1);
			mt.start = NONE; // filled by the head below for linear unitigs
			meta[u].len = mt.len, meta[u].circ = mt.circ, meta[u].n = mt.n, meta[u].first = mt.first, meta[u].end = mt.end;
			if (cyc) meta[u].start = NONE;
		}
		if (a.rk[v] == 0 && !cyc) meta[u].start = v;
	}
}

// unitig-graph arcs (asm.c:181-202): an arc u->v of the read graph joins two unitigs when u^1 and v are unitig ends
__global__ void k_ug_mark(const DUtgMeta *meta, uint32_t n_utg, int32_t *mark)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_utg; i += gridDim.x * blockDim.x) {
		if (meta[i].circ) continue;
		// The sequential loop (asm.c:182-186) lets the last write win when walks on a non-symmetric graph share an end
		// vertex, or when start == end; "last" = larger unitig index, end after start = the larger value: atomicMax.
		atomicMax(&mark[meta[i].start], (int32_t)(i << 1 | 0));
		atomicMax(&mark[meta[i].end], (int32_t)(i << 1 | 1));
	}
}

__global__ void k_ug_arcs(const DArc *arc, uint32_t n_arc, const int32_t *mark, const DUtgMeta *meta, DArc *out, uint8_t *flag)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_arc; i += gridDim.x * blockDim.x) {
		const DArc p = arc[i];
		bool emit = false;
		if (!(p.ol_del & MAB_DEL_BIT)) {
			const int32_t mu = mark[(uint32_t)(p.ul >> 32) ^ 1], mv = mark[p.v];
			if (mu >= 0 && mv >= 0) {
				const uint32_t u = (uint32_t)mu ^ 1, ol = p.ol_del & 0x7fffffffu;
				int l = (int)(meta[u >> 1].len - ol);
				if (l < 0) l = 1;
				DArc q;
				q.ul = (uint64_t)u << 32 | (uint32_t)l, q.v = (uint32_t)mv, q.ol_del = ol;
				out[i] = q;
				emit = true;
			}
		}
		flag[i] = emit;
	}
}

__global__ void k_ug_seq(const DUtgMeta *meta, uint32_t n_utg, uint32_t *seq)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_utg; i += gridDim.x * blockDim.x) seq[i] = meta[i].len & 0x7fffffffu;
}

__global__ void k_ug_cycflag(uint32_t n_vtx, const uint32_t *jb_final, uint32_t *is_cyc)
{
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vtx; v += gridDim.x * blockDim.x) is_cyc[v] = jb_final[v] != NONE;
}

// Literal replay of asm.c:121-178 by one thread: used when F and B are not mutually inverse.
__global__ void k_ug_literal(GV g, int32_t *mark, uint64_t *items, uint64_t cap, uint64_t *tmp, DUtgMeta *meta, unsigned long long *out_counts)
{
	if (blockIdx.x || threadIdx.x) return;
	uint32_t n_utg = 0;
	uint64_t n_items = 0;
	out_counts[2] = 0;
	for (uint32_t v = 0; v < g.n_vtx; ++v) {
		if (n_items + 2ull * g.n_vtx + 2 > cap) { out_counts[2] = 1; break; } // overlapping walks outgrew the buffer
		if ((g.seq[v >> 1] & MAB_DEL_BIT) || (uint32_t)g.idx[v] == 0 || mark[v]) continue;
		mark[v] = 1;
		uint32_t start = v, end = v ^ 1, len = 0, w = v, x, l;
		uint64_t nf = 0, nb = 0; // forward items go to items[n_items..], backward items to tmp[] (reversed later)
		while (1) {
			if ((uint32_t)g.idx[w] != 1) break;
			x = g.arc[g.idx[w] >> 32].v;
			if ((uint32_t)g.idx[x ^ 1] != 1) break;
			mark[x] = mark[w ^ 1] = 1;
			l = (uint32_t)g.arc[g.idx[w] >> 32].ul;
			items[n_items + nf++] = (uint64_t)w << 32 | l;
			end = x ^ 1, len += l;
			w = x;
			if (x == v) break;
		}
		bool circ = false;
		if (start != (end ^ 1) || nf == 0) {
			l = g.seq[end >> 1] & 0x7fffffffu;
			items[n_items + nf++] = (uint64_t)(end ^ 1) << 32 | l;
			len += l;
			x = v;
			while (1) {
				if ((uint32_t)g.idx[x ^ 1] != 1) break;
				w = g.arc[g.idx[x ^ 1] >> 32].v ^ 1;
				if ((uint32_t)g.idx[w] != 1) break;
				mark[x] = mark[w ^ 1] = 1;
				l = (uint32_t)g.arc[g.idx[w] >> 32].ul;
				tmp[nb++] = (uint64_t)w << 32 | l;
				start = w, len += l;
				x = w;
			}
		} else circ = true, start = end = NONE;
		if (start != NONE) mark[start] = mark[end] = 1;
		if (nb) { // prepend the backward items in walk-reversed order
			for (uint64_t k = nf; k-- > 0;) items[n_items + nb + k] = items[n_items + k];
			for (uint64_t k = 0; k < nb; ++k) items[n_items + k] = tmp[nb - 1 - k];
		}
		DUtgMeta mt;
		mt.len = len & 0x7fffffffu, mt.circ = circ, mt.start = start, mt.end = end, mt.n = (uint32_t)(nf + nb), mt.first = (uint32_t)n_items;
		meta[n_utg++] = mt;
		n_items += nf + nb;
	}
	out_counts[0] = n_utg, out_counts[1] = n_items;
}

void dg_ug_free(MabDev &d, DUnitigs &ug)
{
	d.free(ug.meta); d.free(ug.items);
	dg_free(d, ug.g);
	ug = DUnitigs();
}

void dg_ug_gen(MabDev &d, const DGraph &g, DUnitigs &ug)
{
	const uint32_t n_vtx = g.n_seq * 2;
	ug = DUnitigs();
	GV gv{g.arc, g.idx, g.seq, n_vtx};
	if (n_vtx) {
		const unsigned grid = mab_grid(n_vtx, 256);
		UgArrays a;
		a.F = mab_alloc<uint32_t>(d, n_vtx); a.B = mab_alloc<uint32_t>(d, n_vtx);
		MAB_LAUNCH(d, k_ug_links, grid, 256, 0, gv, a.F, a.B);
		d.zero_scal(SC_COUNT);
		MAB_LAUNCH(d, k_ug_consistent, grid, 256, 0, n_vtx, a.F, a.B, d.d_scal + SC_COUNT);
		const bool consistent = d.get_scal(SC_COUNT) == 0;
		if (consistent) {
			uint32_t *buf[12];
			for (int i = 0; i < 12; ++i) buf[i] = mab_alloc<uint32_t>(d, n_vtx);
			uint64_t *ps_buf0 = mab_alloc<uint64_t>(d, n_vtx), *ps_buf1 = mab_alloc<uint64_t>(d, n_vtx);
			uint64_t *ps = ps_buf0, *ps2 = ps_buf1;
			a.jf = buf[0], a.jb = buf[1], a.mf = buf[2], a.mb = buf[3];
			uint32_t *jf2 = buf[4], *jb2 = buf[5], *mf2 = buf[6], *mb2 = buf[7];
			a.rk = buf[8], a.cnt = buf[9], a.seedmin = buf[10];
			uint32_t *is_cyc = buf[11];
			a.ps = ps;
			MAB_LAUNCH(d, k_ug_init, grid, 256, 0, gv, a);
			int steps = 1; while ((1ull << steps) < (uint64_t)n_vtx + 1) ++steps;
			++steps;
			for (int s = 0; s < steps; ++s) {
				MAB_LAUNCH(d, k_ug_min_step, grid, 256, 0, n_vtx, a.jf, a.jb, a.mf, a.mb, jf2, jb2, mf2, mb2);
				uint32_t *t;
				t = a.jf, a.jf = jf2, jf2 = t; t = a.jb, a.jb = jb2, jb2 = t;
				t = a.mf, a.mf = mf2, mf2 = t; t = a.mb, a.mb = mb2, mb2 = t;
			}
			// a.jb now: NONE for path vertices, still live for cycle vertices
			MAB_LAUNCH(d, k_ug_cycflag, grid, 256, 0, n_vtx, a.jb, is_cyc);
			uint32_t *jb_final = jb2; // reuse: copy the final jb aside because k_ug_cut overwrites a.jb
			MAB_CUDA(cudaMemcpyAsync(jb_final, a.jb, (size_t)n_vtx * 4, cudaMemcpyDeviceToDevice, d.stream));
			MAB_LAUNCH(d, k_ug_cut, grid, 256, 0, gv, a, jb_final);
			uint32_t *rk2 = mf2, *cnt2 = mb2; // the min buffers are free now (values kept in seedmin)
			for (int s = 0; s < steps; ++s) {
				MAB_LAUNCH(d, k_ug_rank_step, grid, 256, 0, n_vtx, a.jb, a.rk, a.ps, a.jf, a.cnt, jb2, rk2, ps2, jf2, cnt2);
				uint32_t *t; uint64_t *t8;
				t = a.jb, a.jb = jb2, jb2 = t; t = a.rk, a.rk = rk2, rk2 = t; t8 = a.ps, a.ps = ps2, ps2 = t8;
				t = a.jf, a.jf = jf2, jf2 = t; t = a.cnt, a.cnt = cnt2, cnt2 = t;
			}
			// seeds -> unitig numbers and item offsets
			uint32_t *flag_seed = jb2, *n_at_seed = jf2, *utg_of = rk2, *first_of = cnt2;
			MAB_LAUNCH(d, k_ug_heads, grid, 256, 0, gv, a, is_cyc, flag_seed, n_at_seed);
			size_t tb = 0, tb2 = 0;
			cub::DeviceScan::ExclusiveSum(nullptr, tb, flag_seed, utg_of, (int)n_vtx, d.stream);
			cub::DeviceScan::ExclusiveSum(nullptr, tb2, n_at_seed, first_of, (int)n_vtx, d.stream);
			void *tmp = d.tmp(tb > tb2 ? tb : tb2);
			cub::DeviceScan::ExclusiveSum(tmp, tb, flag_seed, utg_of, (int)n_vtx, d.stream);
			cub::DeviceScan::ExclusiveSum(tmp, tb2, n_at_seed, first_of, (int)n_vtx, d.stream);
			d.n_lib += 2;
			uint32_t last[4];
			MAB_CUDA(cudaMemcpyAsync(&last[0], flag_seed + n_vtx - 1, 4, cudaMemcpyDeviceToHost, d.stream));
			MAB_CUDA(cudaMemcpyAsync(&last[1], utg_of + n_vtx - 1, 4, cudaMemcpyDeviceToHost, d.stream));
			MAB_CUDA(cudaMemcpyAsync(&last[2], n_at_seed + n_vtx - 1, 4, cudaMemcpyDeviceToHost, d.stream));
			MAB_CUDA(cudaMemcpyAsync(&last[3], first_of + n_vtx - 1, 4, cudaMemcpyDeviceToHost, d.stream));
			d.sync();
			ug.n_utg = last[0] + last[1];
			ug.n_items = (uint64_t)last[2] + last[3];
			ug.meta = mab_alloc<DUtgMeta>(d, ug.n_utg);
			ug.items = mab_alloc<uint64_t>(d, ug.n_items);
			if (ug.n_utg) MAB_LAUNCH(d, k_ug_emit, grid, 256, 0, gv, a, is_cyc, utg_of, first_of, ug.items, ug.meta);
			for (int i = 0; i < 12; ++i) d.free(buf[i]);
			d.free(ps_buf0); d.free(ps_buf1);
		} else {
			int32_t *mark = mab_alloc<int32_t>(d, n_vtx);
			MAB_CUDA(cudaMemsetAsync(mark, 0, (size_t)n_vtx * 4, d.stream));
			ug.meta = mab_alloc<DUtgMeta>(d, n_vtx);
			const uint64_t cap = (uint64_t)n_vtx * 6 + 8;
			ug.items = mab_alloc<uint64_t>(d, cap);
			uint64_t *tmp = mab_alloc<uint64_t>(d, (size_t)n_vtx + 1);
			MAB_LAUNCH(d, k_ug_literal, 1, 32, 0, gv, mark, ug.items, cap, tmp, ug.meta, d.d_scal + SC_TMP0);
			ug.n_utg = (uint32_t)d.get_scal(SC_TMP0);
			ug.n_items = d.h_scal[SC_TMP0 + 1];
			if (d.h_scal[SC_TMP0 + 2]) { fprintf(stderr, "[E::miniasm_b200] ma_ug_gen: unitig walks on a non-symmetric graph outgrew the item buffer\n"); exit(76); }
			d.free(mark); d.free(tmp);
		}
		d.free(a.F); d.free(a.B);
	}
	// unitig graph
	DGraph &q = ug.g;
	dg_set_nseq(d, q, ug.n_utg);
	dg_reserve(d, q, g.n_arc ? g.n_arc : 1);
	q.n_arc = 0, q.is_srt = false, q.is_symm = false, q.len_bits = 32;
	if (ug.n_utg) {
		MAB_LAUNCH(d, k_ug_seq, mab_grid(ug.n_utg, 256), 256, 0, ug.meta, ug.n_utg, q.seq);
		if (g.n_arc) {
			int32_t *mark = mab_alloc<int32_t>(d, n_vtx);
			uint8_t *flag = mab_alloc<uint8_t>(d, g.n_arc);
			MAB_CUDA(cudaMemsetAsync(mark, 0xff, (size_t)n_vtx * 4, d.stream));
			MAB_LAUNCH(d, k_ug_mark, mab_grid(ug.n_utg, 256), 256, 0, ug.meta, ug.n_utg, mark);
			MAB_LAUNCH(d, k_ug_arcs, mab_grid(g.n_arc, 256), 256, 0, g.arc, g.n_arc, mark, ug.meta, q.arc2, flag);
			size_t tb = 0;
			unsigned long long *d_n = d.d_scal + SC_NSEL;
			cub::DeviceSelect::Flagged(nullptr, tb, q.arc2, flag, q.arc, d_n, (int)g.n_arc, d.stream);
			void *tmp = d.tmp(tb);
			cub::DeviceSelect::Flagged(tmp, tb, q.arc2, flag, q.arc, d_n, (int)g.n_arc, d.stream);
			++d.n_lib;
			q.n_arc = (uint32_t)d.get_scal(SC_NSEL);
			d.free(mark); d.free(flag);
		}
	}
	dg_cleanup(d, q);
}
