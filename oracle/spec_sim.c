/* spec_sim.c -- CPU model of the speculative prefix-commit rounds that clean_dev.cu runs for asg_pop_bubble
 * (test infrastructure, like everything under oracle/: never linked into or called by the product).
 *
 * It replays, on the host and with the oracle port's own bubble walk (pop_one of ma_oracle.c split into "walk" and
 * "apply"), exactly what the CUDA rounds do -- candidates stamp their walk set with atomicMin(tag[read], v0); every
 * source whose walk set carries a smaller stamp is invalid; x* = smallest invalid vertex; candidates below x* commit --
 * and compares the final deletion state with the sequential pass.  It exists to check two variations before they are
 * trusted on the GPU: dropping the own-cell validity check of NON-sources (MAB_BUB_SKIP_OWN) and evaluating a sliding
 * window of vertices per round (MAB_SPEC_WINDOW), and to count the rounds each variant needs.
 *
 * usage: spec_sim in.paf      -> one line per variant: rounds, pops, equal-to-sequential (1/0)                          */
#include "ma_oracle.c"

typedef struct { uint8_t *seq_del, *arc_del; } state_t;

static state_t snapshot(const asg_t *g)
{
	state_t s;
	uint32_t i;
	s.seq_del = (uint8_t*)malloc(g->n_seq ? g->n_seq : 1);
	s.arc_del = (uint8_t*)malloc(g->n_arc ? g->n_arc : 1);
	for (i = 0; i < g->n_seq; ++i) s.seq_del[i] = g->seq[i].del;
	for (i = 0; i < g->n_arc; ++i) s.arc_del[i] = g->arc[i].del;
	return s;
}
static void restore(asg_t *g, const state_t *s)
{
	uint32_t i;
	for (i = 0; i < g->n_seq; ++i) g->seq[i].del = s->seq_del[i];
	for (i = 0; i < g->n_arc; ++i) g->arc[i].del = s->arc_del[i];
}
static int same(const asg_t *g, const state_t *s)
{
	uint32_t i;
	for (i = 0; i < g->n_seq; ++i) if (g->seq[i].del != s->seq_del[i]) return 0;
	for (i = 0; i < g->n_arc; ++i) if (g->arc[i].del != s->arc_del[i]) return 0;
	return 1;
}

static int is_source(const asg_t *g, uint32_t v)
{
	uint32_t i, live = 0, nv = A_N(g, v);
	const asg_arc_t *av = A_A(g, v);
	if (nv < 2 || g->seq[v >> 1].del) return 0;
	for (i = 0; i < nv; ++i) live += !av[i].del;
	return live > 1;
}

/* pop_one on a scratch copy of the flags it would change: returns 1 if v0 pops, and the walk set (b list) */
static uint32_t g_sink; /* sink of the last successful walk */
/* the traversal of pop_one (asg.c:360-409) without the backtrack: 1 if v0 would pop; *visited = its b list, g_sink = sink */
static int walk_only(asg_t *g, uint32_t v0, int max_dist, binfo_t *a, u32v *S, u32v *T, u32v *b, u32v *e, u32v *visited)
{
	uint32_t i, pending = 0;
	int ret = 0;
	visited->n = 0;
	if (g->seq[v0 >> 1].del || A_N(g, v0) < 2) return 0;
	S->n = T->n = b->n = e->n = 0;
	a[v0].c = a[v0].d = 0;
	vpush(S, v0);
	do {
		uint32_t v = S->a[--S->n], d = a[v].d, c = a[v].c, nv = A_N(g, v);
		const asg_arc_t *av = A_A(g, v);
		for (i = 0; i < nv; ++i) {
			uint32_t w = av[i].v, l = (uint32_t)av[i].ul;
			binfo_t *t = &a[w];
			if (w == v0) goto reset;
			if (av[i].del) continue;
			vpush(e, (uint32_t)(g->idx[v] >> 32) + i);
			if (d + l > (uint32_t)max_dist) break;
			if (!t->s) {
				uint32_t k, nx = A_N(g, w ^ 1);
				const asg_arc_t *ax = A_A(g, w ^ 1);
				vpush(b, w);
				t->p = v; t->s = 1; t->d = d + l; t->r = 0;
				for (k = 0; k < nx; ++k) t->r += !ax[k].del;
				++pending;
			} else {
				if (c + 1 > t->c || (c + 1 == t->c && d + l > t->d)) t->p = v;
				if (c + 1 > t->c) t->c = c + 1;
				if (d + l < t->d) t->d = d + l;
			}
			t->r = (t->r - 1) & 0x7fffffffu;
			if (t->r == 0) { if (A_N(g, w)) vpush(S, w); else vpush(T, w); --pending; }
		}
		if (i < nv || S->n == 0) goto reset;
	} while (S->n > 1 || pending);
	ret = 1, g_sink = S->a[0];
reset:
	/* like the GPU rounds, the walk set of a failed traversal is what it touched before giving up */
	for (i = 0; i < b->n; ++i) { binfo_t *t = &a[b->a[i]]; vpush(visited, b->a[i]); t->s = t->c = t->d = 0; }
	return ret;
}

/* excuse = 1: an invalid source u does not stop the prefix if its own read was stamped by a VALID candidate A < u and
 * u is neither A^1 nor A's sink: every other vertex on a read of A's walk set has all its live out-arcs (or, on the
 * complement strand, the complements of all its live in-arcs) inside A's region, so after A's pop at most the restored
 * path arc is left -- u stops being a source, and in this deletion-only pass it stays a no-op for good. */
static uint32_t spec_rounds(asg_t *g, int max_dist, int skip_own, int window, int excuse, uint64_t *n_pop)
{
	uint32_t n_vtx = g->n_seq * 2, lo = 0, rounds = 0, W = n_vtx, v, i;
	binfo_t *a = (binfo_t*)calloc(n_vtx ? n_vtx : 1, sizeof(binfo_t));
	u32v S = {0,0,0}, T = {0,0,0}, b = {0,0,0}, e = {0,0,0}, vis = {0,0,0};
	uint32_t *tag = (uint32_t*)malloc(4 * (size_t)(g->n_seq ? g->n_seq : 1));
	uint8_t *cand = (uint8_t*)calloc(n_vtx ? n_vtx : 1, 1);
	uint32_t *sink = (uint32_t*)malloc(4 * (size_t)(n_vtx ? n_vtx : 1)), *mn = (uint32_t*)malloc(4 * (size_t)(n_vtx ? n_vtx : 1));
	*n_pop = 0;
	while (lo < n_vtx) {
		uint32_t end = (!window || (uint64_t)lo + W >= n_vtx) ? n_vtx : lo + W, xs = 0xffffffffu, hi, n_cand = 0, n_src = 0;
		memset(tag, 0xff, 4 * (size_t)g->n_seq);
		for (v = lo; v < end; ++v) {                                   /* k_bub_sources + k_bub_eval */
			cand[v] = 0;
			if (!is_source(g, v)) continue;
			++n_src;
			if (walk_only(g, v, max_dist, a, &S, &T, &b, &e, &vis)) {
				cand[v] = 1, ++n_cand, sink[v] = g_sink;
				if (v < tag[v >> 1]) tag[v >> 1] = v;
				for (i = 0; i < vis.n; ++i) if (v < tag[vis.a[i] >> 1]) tag[vis.a[i] >> 1] = v;
			}
		}
		if (n_src == 0 || n_cand == 0) {
			if (end == n_vtx) break;
			W = (uint64_t)W * 4 >= n_vtx ? n_vtx : W * 4;
			lo = end;
			continue;
		}
		if (!skip_own) for (v = lo; v < end; ++v) if (tag[v >> 1] < v && v < xs) xs = v;   /* k_bub_check_own */
		for (v = lo; v < end; ++v) {                                   /* k_bub_check_walk: smallest stamp on the walk set */
			uint32_t m;
			mn[v] = 0xffffffffu;
			if (!is_source(g, v)) continue;
			walk_only(g, v, max_dist, a, &S, &T, &b, &e, &vis);
			m = tag[v >> 1];
			for (i = 0; i < vis.n; ++i) if (tag[vis.a[i] >> 1] < m) m = tag[vis.a[i] >> 1];
			mn[v] = m;
		}
		for (v = lo; v < end; ++v) {
			uint32_t A;
			if (!is_source(g, v) || mn[v] >= v) continue;              /* valid */
			A = tag[v >> 1];
			if (excuse && A < v && A >= lo && cand[A] && mn[A] >= A && v != (A ^ 1) && v != sink[A]) continue; /* becomes a non-source */
			if (v < xs) xs = v;
		}
		hi = xs >= end ? end : xs;
		for (v = lo; v < hi; ++v)                                      /* k_bub_commit: all of them evaluated on the round's state ... */
			if (cand[v] && mn[v] >= v) *n_pop += pop_one(g, v, max_dist, a, &S, &T, &b, &e) & 1;   /* (valid ones: they are cell-disjoint, so applying in turn is the same) */
		if (window) {
			if (hi >= end) W = (uint64_t)W * 4 >= n_vtx ? n_vtx : W * 4;
			else { uint64_t w = 16ull * (hi - lo + 1); W = w < 4096 ? 4096 : (w < W ? (uint32_t)w : W); }
		}
		lo = hi;
		++rounds;
	}
	free(a); free(S.a); free(T.a); free(b.a); free(e.a); free(vis.a); free(tag); free(cand); free(sink); free(mn);
	return rounds;
}

/* ------------------------------------------------------------------------------------------------------------
 * The same for asg_cut_tip (TipRule of clean_dev.cu): a candidate stamps every cell it reads (utg_end of v and of
 * every chain vertex: the vertex's read and its single neighbour's) or writes (read_del of each chain read: the
 * read and every read one of its arcs points to).  excuse = 1: an invalid vertex whose read a VALID smaller candidate
 * is going to delete does not stop the prefix -- on a deleted read the rule returns at its first test, for good.
 * ------------------------------------------------------------------------------------------------------------ */
static int utg_end_cells(const asg_t *g, uint32_t v, uint64_t *lw, u32v *cells)
{
	const asg_arc_t *av = A_A(g, v ^ 1), *aw;
	uint32_t i, n = 0, last = 0, nv = A_N(g, v ^ 1), w, nw;
	vpush(cells, v >> 1);
	for (i = 0; i < nv; ++i) if (!av[i].del) last = i, ++n;
	if (n == 0) return E_TIP;
	if (n > 1) return E_MOUT;
	if (lw) *lw = av[last].ul << 32 | av[last].v;
	w = av[last].v ^ 1; aw = A_A(g, w); nw = A_N(g, w);
	vpush(cells, w >> 1);
	for (i = n = 0; i < nw; ++i) if (!aw[i].del) ++n;
	return n == 1 ? E_MERGE : E_MNEI;
}

/* 1 if v cuts a tip on the current state; cells = read set (+ write set when it acts), chain = the reads it deletes */
static int tip_eval(const asg_t *g, uint32_t v, int max_ext, u32v *cells, u32v *chain)
{
	int r;
	uint64_t lw = 0;
	uint32_t x = v, i, k;
	cells->n = chain->n = 0;
	if (g->seq[v >> 1].del) { vpush(cells, v >> 1); return 0; }
	if (utg_end_cells(g, v, 0, cells) != E_TIP) return 0;
	vpush(chain, x >> 1);
	do {
		r = utg_end_cells(g, x ^ 1, &lw, cells);
		if (r != E_MERGE) break;
		x = (uint32_t)lw; vpush(chain, x >> 1);
	} while (--max_ext > 0);
	if (r == E_MERGE) return 0;
	for (i = 0; i < chain->n; ++i)                       /* write set: seq_del_cells */
		for (vpush(cells, chain->a[i]), k = 0; k < 2; ++k) {
			uint32_t u = chain->a[i] << 1 | k, j, nu = A_N(g, u);
			const asg_arc_t *au = A_A(g, u);
			for (j = 0; j < nu; ++j) vpush(cells, au[j].v >> 1);
		}
	return 1;
}

static uint32_t tip_rounds(asg_t *g, int max_ext, int window, int excuse, uint32_t *n_cut)
{
	uint32_t n_vtx = g->n_seq * 2, lo = 0, rounds = 0, W = n_vtx, v, i;
	u32v cells = {0,0,0}, chain = {0,0,0};
	uint32_t *tag = (uint32_t*)malloc(4 * (size_t)(g->n_seq + 1)), *dtag = (uint32_t*)malloc(4 * (size_t)(g->n_seq + 1));
	uint8_t *cand = (uint8_t*)calloc(n_vtx + 1, 1), *valid = (uint8_t*)calloc(n_vtx + 1, 1);
	*n_cut = 0;
	while (lo < n_vtx) {
		uint32_t end = (!window || (uint64_t)lo + W >= n_vtx) ? n_vtx : lo + W, xs = 0xffffffffu, hi, n_cand = 0;
		memset(tag, 0xff, 4 * (size_t)g->n_seq); memset(dtag, 0xff, 4 * (size_t)g->n_seq);
		for (v = lo; v < end; ++v) {                                   /* k_spec_eval */
			cand[v] = (uint8_t)tip_eval(g, v, max_ext, &cells, &chain);
			if (!cand[v]) continue;
			++n_cand;
			for (i = 0; i < cells.n; ++i) if (v < tag[cells.a[i]]) tag[cells.a[i]] = v;
			for (i = 0; i < chain.n; ++i) if (v < dtag[chain.a[i]]) dtag[chain.a[i]] = v;
		}
		if (n_cand == 0) { if (end == n_vtx) break; W = (uint64_t)W * 4 >= n_vtx ? n_vtx : W * 4; lo = end; continue; }
		for (v = lo; v < end; ++v) {                                   /* k_spec_check: smallest stamp on what v reads (+ writes) */
			uint32_t m = 0xffffffffu;
			tip_eval(g, v, max_ext, &cells, &chain);
			for (i = 0; i < cells.n; ++i) if (tag[cells.a[i]] < m) m = tag[cells.a[i]];
			valid[v] = m >= v;
		}
		for (v = lo; v < end; ++v) {
			uint32_t A;
			if (valid[v]) continue;
			A = dtag[v >> 1];
			if (excuse && A < v && cand[A] && valid[A]) continue;       /* its read is about to be deleted: a no-op from then on */
			if (v < xs) xs = v;
		}
		hi = xs >= end ? end : xs;
		for (v = lo; v < hi; ++v)                                      /* k_spec_commit (valid candidates only) */
			if (cand[v] && valid[v] && tip_eval(g, v, max_ext, &cells, &chain)) {
				u32v del = {0,0,0};
				for (i = 0; i < chain.n; ++i) vpush(&del, chain.a[i]);
				for (i = 0; i < del.n; ++i) read_del(g, del.a[i]);
				free(del.a);
				++*n_cut;
			}
		if (window) {
			if (hi >= end) W = (uint64_t)W * 4 >= n_vtx ? n_vtx : W * 4;
			else { uint64_t w = 16ull * (hi - lo + 1); W = w < 4096 ? 4096 : (w < W ? (uint32_t)w : W); }
		}
		lo = hi;
		++rounds;
	}
	free(cells.a); free(chain.a); free(tag); free(dtag); free(cand); free(valid);
	return rounds;
}

int main(int argc, char *argv[])
{
	ma_opt_t opt;
	sdict_t *d;
	ma_hit_t *hit;
	ma_sub_t *sub, *sub2;
	size_t n_hits;
	asg_t *g;
	float cov;
	int k;
	if (argc < 2) { fprintf(stderr, "usage: spec_sim in.paf\n"); return 1; }
	ma_verbose = 0;
	ma_opt_init(&opt);
	d = sd_init();
	hit = ma_hit_read(argv[1], opt.min_span, opt.min_match, d, &n_hits, 1, 0);
	sub = ma_hit_sub(opt.min_dp, opt.min_iden, 0, n_hits, hit, d->n_seq);
	n_hits = ma_hit_cut(sub, opt.min_span, n_hits, hit);
	n_hits = ma_hit_flt(sub, (int)(opt.max_hang * 1.5), (int)(opt.min_ovlp * .5), n_hits, hit, &cov);
	sub2 = ma_hit_sub(opt.min_dp, opt.min_iden, opt.min_span / 2, n_hits, hit, d->n_seq);
	n_hits = ma_hit_cut(sub2, opt.min_span, n_hits, hit);
	ma_sub_merge(d->n_seq, sub, sub2);
	free(sub2);
	n_hits = ma_hit_contained(&opt, d, sub, n_hits, hit);
	g = ma_sg_gen(&opt, d, sub, n_hits, hit);
	asg_arc_del_trans(g, opt.gap_fuzz);
	{ /* tips, on the state main.c:161 sees them */
		state_t before = snapshot(g), seq;
		uint32_t v, n_vtx = g->n_seq * 2, n_seq_cut = 0, i, np, *path = (uint32_t*)malloc(4 * ((size_t)opt.max_ext + 2));
		for (v = 0; v < n_vtx; ++v) { /* cut_generic(E_TIP, E_MERGE, negate) without the cleanup */
			if (g->seq[v >> 1].del || utg_end(g, v, 0) != E_TIP) continue;
			if (walk(g, v, opt.max_ext, path, &np) == E_MERGE) continue;
			for (i = 0; i < np; ++i) read_del(g, path[i] >> 1);
			++n_seq_cut;
		}
		free(path);
		seq = snapshot(g);
		printf("tips sequential: %u cut\n", n_seq_cut);
		for (k = 0; k < 4; ++k) {
			uint32_t n_cut, rounds;
			restore(g, &before);
			rounds = tip_rounds(g, opt.max_ext, k & 1, k >> 1, &n_cut);
			printf("tips window=%d excuse=%d: rounds %u cut %u equal %d\n", k & 1, k >> 1, rounds, n_cut, same(g, &seq));
		}
		restore(g, &before);
		free(before.seq_del); free(before.arc_del); free(seq.seq_del); free(seq.arc_del);
	}
	asg_cut_tip(g, opt.max_ext);
	if (!g->is_symm) asg_symm(g);
	{
		state_t before = snapshot(g), seq;
		binfo_t *a = (binfo_t*)calloc(g->n_seq * 2 + 1, sizeof(binfo_t));
		u32v S = {0,0,0}, T = {0,0,0}, b = {0,0,0}, e = {0,0,0};
		uint32_t v, n_vtx = g->n_seq * 2;
		uint64_t n_seq_pop = 0;
		for (v = 0; v < n_vtx; ++v) if (is_source(g, v)) n_seq_pop += pop_one(g, v, opt.bub_dist, a, &S, &T, &b, &e) & 1; /* asg_pop_bubble without the cleanup */
		seq = snapshot(g);
		printf("sequential: %lu pops (%u vertices, %u arcs)\n", (unsigned long)n_seq_pop, n_vtx, g->n_arc);
		for (k = 0; k < 8; ++k) {
			uint64_t n_pop;
			uint32_t rounds;
			restore(g, &before);
			rounds = spec_rounds(g, opt.bub_dist, k & 1, k >> 1 & 1, k >> 2, &n_pop);
			printf("skip_own=%d window=%d excuse=%d: rounds %u pops %lu equal %d\n", k & 1, k >> 1 & 1, k >> 2, rounds, (unsigned long)n_pop, same(g, &seq));
		}
		free(a);
	}
	return 0;
}
