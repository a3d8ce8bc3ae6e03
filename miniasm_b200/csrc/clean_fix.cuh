// clean_fix.cuh -- the order-dependent cleaning passes of stage (iii) as a TIMESTAMP FIXED POINT.
//   asg_cut_tip asg.c:238-254 | asg_cut_internal asg.c:256-272 | asg_cut_biloop asg.c:274-306 |
//   asg_pop_bubble asg.c:412-433 (asg_bub_pop1 asg.c:360-409, asg_bub_backtrack asg.c:338-357)
//
// The reference runs `for v ascending: decide on the current deletion bits, then delete`.  Iteration v sees the
// deletions of every u < v, so a decide-on-a-snapshot parallel-for is wrong.  All four passes are
// DELETION-ONLY as far as a later iteration can tell (the bubble backtrack deletes a region and revives its best path
// inside one iteration; the revived bits were live before).  The whole pass is therefore described by one number per
// bit: T(bit) = 0 if it was deleted before the pass, u+1 if iteration u deletes it, LIVE otherwise, and iteration v
// sees "deleted iff T <= v".  T is the unique fixed point of
//       T = init  min  { v+1 on every bit that v deletes when it decides under the view T }      (*)
// (unique by induction on v: the view of v only holds stamps of u < v).  Jacobi iteration of (*) -- every vertex
// decides in parallel under T_old, stamps T_new with atomicMin, repeat until T_new == T_old -- makes the decisions of
// at least one more vertex final per sweep in the worst case and, on real graphs, of everything whose dependency CHAIN
// is one link longer: a handful of sweeps instead of one round per bubble.  Decisions are pure functions of T_old, so a
// sweep has no races and no ordering; counters (tips cut, bubbles popped) are those of the last sweep, which ran on
// the fixed point itself.
//
// This header is the whole algorithm, written once for the device and for the host (the host build is test
// infrastructure: tests/hostsim/fix_host.cpp runs the same sweeps sequentially against the reference on the CPU tier).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FX_HD __host__ __device__ __forceinline__
#define FX_HD_NOINL __host__ __device__
#define FX_MEMBER __host__ __device__ __forceinline__
#else
#define FX_HD static inline
#define FX_HD_NOINL static
#define FX_MEMBER inline
#endif

#ifndef MAB_DEL_BIT
#define MAB_DEL_BIT 0x80000000u
#endif

#define FX_ET_MERGEABLE 0
#define FX_ET_TIP       1
#define FX_ET_MULTI_OUT 2
#define FX_ET_MULTI_NEI 3

constexpr uint32_t FX_LIVE = 0xffffffffu;

FX_HD void fx_stamp(uint32_t *p, uint32_t t)
{
#if defined(__CUDA_ARCH__)
	if (*p > t) atomicMin(p, t);
#else
	if (*p > t) *p = t;
#endif
}

// The graph as a sweep sees it: T_old to read, T_new to stamp.
struct FxView {
	const DArc *arc;            // structure: never written during the sweeps
	const uint64_t *idx;
	const uint32_t *ts, *ta;    // T_old: per read / per arc
	uint32_t *ns, *na;          // T_new
	uint32_t n_vtx;
	FX_MEMBER uint32_t at(uint32_t i) const { return ta[i]; }
	FX_MEMBER uint32_t st(uint32_t s) const { return ts[s]; }
	FX_MEMBER void stamp_a(uint32_t i, uint32_t t) const { fx_stamp(&na[i], t); }
	FX_MEMBER void stamp_s(uint32_t s, uint32_t t) const { fx_stamp(&ns[s], t); }
};

// The same interface straight on the deletion bits, stamping nothing: the PROBE sweep that opens every pass.  T_init is
// "0 where the bit is set, LIVE elsewhere", so deciding on the bits IS the first Jacobi sweep; if nobody acts in it the
// pass is over (the reference's loop would not have changed anything either) and no timestamp array was ever touched.
struct FxProbe {
	const DArc *arc;
	const uint64_t *idx;
	const uint32_t *seq;
	uint32_t n_vtx;
	FX_MEMBER uint32_t at(uint32_t i) const { return arc[i].ol_del & MAB_DEL_BIT ? 0u : FX_LIVE; }
	FX_MEMBER uint32_t st(uint32_t s) const { return seq[s] & MAB_DEL_BIT ? 0u : FX_LIVE; }
	FX_MEMBER void stamp_a(uint32_t, uint32_t) const {}
	FX_MEMBER void stamp_s(uint32_t, uint32_t) const {}
};

// asg_is_utg_end (asg.c:204-222) as iteration `me` sees it
template <class V> FX_HD int fx_utg_end(const V &g, uint32_t me, uint32_t v, uint64_t *lw)
{
	const uint64_t iv = g.idx[v ^ 1];
	const uint32_t nv0 = (uint32_t)iv, off = (uint32_t)(iv >> 32);
	uint32_t nv = 0, i0 = 0;
	for (uint32_t i = 0; i < nv0; ++i)
		if (g.at(off + i) > me) i0 = off + i, ++nv;
	if (nv == 0) return FX_ET_TIP;
	if (nv > 1) return FX_ET_MULTI_OUT;
	const DArc a = g.arc[i0];
	if (lw) *lw = a.ul << 32 | a.v;
	const uint64_t iw = g.idx[a.v ^ 1];
	const uint32_t nw0 = (uint32_t)iw, offw = (uint32_t)(iw >> 32);
	uint32_t nw = 0;
	for (uint32_t i = 0; i < nw0; ++i)
		if (g.at(offw + i) > me) ++nw;
	return nw != 1 ? FX_ET_MULTI_NEI : FX_ET_MERGEABLE;
}

// asg_extend (asg.c:224-236) without materialising the path: `chain(vertex)` sees every vertex the reference pushes
template <class V, class ChainFn>
FX_HD int fx_extend(const V &g, uint32_t me, uint32_t v, int max_ext, ChainFn chain, uint32_t *last)
{
	int ret;
	uint64_t lw = 0;
	chain(v);
	*last = v;
	do {
		ret = fx_utg_end(g, me, v ^ 1, &lw);
		if (ret != 0) break;
		v = (uint32_t)lw;
		chain(v);
		*last = v;
	} while (--max_ext > 0);
	return ret;
}

// asg_arc_del(g, v, w, 1) (asg.h:55-61): every arc v->w
template <class V> FX_HD void fx_arc_del(const V &g, uint32_t v, uint32_t w, uint32_t t)
{
	const uint64_t iv = g.idx[v];
	const uint32_t off = (uint32_t)(iv >> 32);
	for (uint32_t i = 0; i < (uint32_t)iv; ++i)
		if (g.arc[off + i].v == w) g.stamp_a(off + i, t);
}

// asg_seq_del (asg.h:64-77): the read, all arcs of both strands and their complements
template <class V> FX_HD void fx_seq_del(const V &g, uint32_t s, uint32_t t)
{
	g.stamp_s(s, t);
	for (uint32_t k = 0; k < 2; ++k) {
		const uint32_t v = s << 1 | k;
		const uint64_t iv = g.idx[v];
		const uint32_t off = (uint32_t)(iv >> 32);
		for (uint32_t i = 0; i < (uint32_t)iv; ++i) {
			g.stamp_a(off + i, t);
			fx_arc_del(g, g.arc[off + i].v ^ 1, v ^ 1, t);
		}
	}
}

// ---- the three short-unitig cutters: act(g, v) decides under T_old and stamps T_new; true if v acts ----
struct FxTip { // asg_cut_tip
	int max_ext;
	template <class V> FX_MEMBER bool act(const V &g, uint32_t v) const
	{
		if (g.st(v >> 1) <= v) return false;
		if (fx_utg_end(g, v, v, nullptr) != FX_ET_TIP) return false;
		uint32_t last;
		if (fx_extend(g, v, v, max_ext, [](uint32_t) {}, &last) == FX_ET_MERGEABLE) return false;
		// the second walk reads T_old again, the stamps go to T_new: same chain as the one the reference collected first
		fx_extend(g, v, v, max_ext, [&](uint32_t x) { fx_seq_del(g, x >> 1, v + 1); }, &last);
		return true;
	}
};

struct FxInternal { // asg_cut_internal
	int max_ext;
	template <class V> FX_MEMBER bool act(const V &g, uint32_t v) const
	{
		if (g.st(v >> 1) <= v) return false;
		if (fx_utg_end(g, v, v, nullptr) != FX_ET_MULTI_NEI) return false;
		uint32_t last;
		if (fx_extend(g, v, v, max_ext, [](uint32_t) {}, &last) != FX_ET_MULTI_NEI) return false;
		fx_extend(g, v, v, max_ext, [&](uint32_t x) { fx_seq_del(g, x >> 1, v + 1); }, &last);
		return true;
	}
};

struct FxBiloop { // asg_cut_biloop: v->...->x', w->v and w->x; drop w->x (and its complement) if it is the weaker one
	int max_ext;
	template <class V> FX_MEMBER bool act(const V &g, uint32_t v) const
	{
		if (g.st(v >> 1) <= v) return false;
		if (fx_utg_end(g, v, v, nullptr) != FX_ET_MULTI_NEI) return false;
		uint32_t last;
		if (fx_extend(g, v, v, max_ext, [](uint32_t) {}, &last) != FX_ET_MULTI_OUT) return false;
		const uint32_t x = last ^ 1;
		uint32_t w = 0xffffffffu, ov = 0, ox = 0;
		{
			const uint64_t iv = g.idx[v ^ 1];
			const uint32_t off = (uint32_t)(iv >> 32);
			for (uint32_t i = 0; i < (uint32_t)iv; ++i)
				if (g.at(off + i) > v) w = g.arc[off + i].v ^ 1;
		}
		if (w == 0xffffffffu) return false; // cannot happen: MULTI_NEI means exactly one live arc (asg.c:288 asserts it)
		const uint64_t iw = g.idx[w];
		const uint32_t offw = (uint32_t)(iw >> 32);
		for (uint32_t i = 0; i < (uint32_t)iw; ++i) {
			if (g.at(offw + i) <= v) continue;
			const DArc a = g.arc[offw + i];
			if (a.v == x) ox = a.ol_del & ~MAB_DEL_BIT;
			if (a.v == v) ov = a.ol_del & ~MAB_DEL_BIT;
		}
		if (ov == 0 && ox == 0) return false;
		if (!(ov > ox)) return false;
		fx_arc_del(g, w, x, v + 1);
		fx_arc_del(g, x ^ 1, w ^ 1, v + 1);
		return true;
	}
};

// ---------------------------------------------------------------------------------------------
// Bubble popping.  asg_bub_pop1 is a bounded Kahn-style traversal from a source v0 with >= 2 live out-arcs; per
// visited vertex it keeps {best parent p, distance d, read count c, pending in-arcs r}.  The reference indexes one
// n_vtx-sized array by vertex; here each traversal owns a small open-addressing table vertex -> {p,d,c,r} in a
// scratch slot, which is the same map restricted to the visited set.
// ---------------------------------------------------------------------------------------------
struct FxSlot { // one traversal's scratch (a slice of the arrays in FxSlots)
	uint32_t *hkey, *hp, *hd, *hc, *hr;  // [hcap]
	uint32_t *b, *bslot, *S;             // [bcap]
	uint32_t *e;                         // [ecap]
	uint32_t bcap, ecap, hmask;
};

constexpr uint32_t FX_EMPTY = 0xffffffffu;
constexpr uint32_t FX_ON_PATH = 0x80000000u; // flag kept in hr[] after a successful walk (every r is 0 by then)

struct FxBubRes { uint32_t nb, ne, nT, sink; };

FX_HD uint32_t fx_hash(uint32_t key, uint32_t hmask) { return (key * 2654435761u) >> 9 & hmask; }

// slot of `key`, or FX_EMPTY if it was never visited
FX_HD uint32_t fx_lookup(const FxSlot &sl, uint32_t key)
{
	uint32_t h = fx_hash(key, sl.hmask);
	while (sl.hkey[h] != key) {
		if (sl.hkey[h] == FX_EMPTY) return FX_EMPTY;
		h = (h + 1) & sl.hmask;
	}
	return h;
}

template <class V> FX_HD bool fx_is_source(const V &g, uint32_t v)
{
	const uint64_t iv = g.idx[v];
	const uint32_t nv = (uint32_t)iv, off = (uint32_t)(iv >> 32);
	if (nv < 2 || g.st(v >> 1) <= v) return false;
	uint32_t live = 0;
	for (uint32_t i = 0; i < nv; ++i) live += g.at(off + i) > v;
	return live > 1;
}

// the traversal of asg_bub_pop1 as iteration v0 sees the graph: 1 = bubble resolved, 0 = nothing to pop, -1 = scratch too small
template <class V> FX_HD_NOINL int fx_bub_walk(const V &g, uint32_t v0, uint32_t max_dist, const FxSlot &sl, FxBubRes *out)
{
	uint32_t nb = 0, ne = 0, nT = 0, nS = 0, n_pending = 0;
	int ret = 0;
	sl.S[nS++] = v0;
	do {
		const uint32_t v = sl.S[--nS];
		uint32_t d = 0, c = 0;
		if (v != v0) { const uint32_t h = fx_lookup(sl, v); d = sl.hd[h], c = sl.hc[h]; }
		const uint64_t iv = g.idx[v];
		const uint32_t nv = (uint32_t)iv, off = (uint32_t)(iv >> 32);
		uint32_t i;
		for (i = 0; i < nv; ++i) {
			const DArc a = g.arc[off + i];
			const uint32_t w = a.v, l = (uint32_t)a.ul;
			if (w == v0) goto done;                            // a cycle through the source (tested before the del bit, asg.c:377)
			if (g.at(off + i) <= v0) continue;
			if (ne == sl.ecap) { ret = -1; goto done; }
			sl.e[ne++] = off + i;
			if (d + l > max_dist) break;                       // too far
			uint32_t h = fx_hash(w, sl.hmask);
			while (sl.hkey[h] != FX_EMPTY && sl.hkey[h] != w) h = (h + 1) & sl.hmask;
			if (sl.hkey[h] == FX_EMPTY) {                      // first visit
				if (nb == sl.bcap) { ret = -1; goto done; }
				sl.hkey[h] = w; sl.bslot[nb] = h; sl.b[nb++] = w;
				sl.hp[h] = v, sl.hd[h] = d + l, sl.hc[h] = 0;
				uint32_t r = 0;                                // count_out(w^1): live arcs only
				const uint64_t ix = g.idx[w ^ 1];
				const uint32_t offx = (uint32_t)(ix >> 32);
				for (uint32_t k = 0; k < (uint32_t)ix; ++k) r += g.at(offx + k) > v0;
				sl.hr[h] = r;
				++n_pending;
			} else {
				if (c + 1 > sl.hc[h] || (c + 1 == sl.hc[h] && d + l > sl.hd[h])) sl.hp[h] = v;
				if (c + 1 > sl.hc[h]) sl.hc[h] = c + 1;
				if (d + l < sl.hd[h]) sl.hd[h] = d + l;
			}
			sl.hr[h] = (sl.hr[h] - 1) & 0x7fffffffu;
			if (sl.hr[h] == 0) {
				if ((uint32_t)g.idx[w]) { if (nS == sl.bcap) { ret = -1; goto done; } sl.S[nS++] = w; }
				else ++nT;                                     // a tip
				--n_pending;
			}
		}
		if (i < nv || nS == 0) goto done;
	} while (nS > 1 || n_pending);
	ret = 1;
	out->sink = sl.S[0];
done:
	out->nb = nb, out->ne = ne, out->nT = nT;
	return ret;
}

FX_HD void fx_bub_reset(const FxSlot &sl, uint32_t nb)
{
	for (uint32_t i = 0; i < nb; ++i) sl.hkey[sl.bslot[i]] = FX_EMPTY;
}

// is s->t one of the arcs asg_bub_backtrack revives (a best-path arc p[x]->x or its complement x'->p[x]')?
FX_HD bool fx_revived(const FxSlot &sl, uint32_t s, uint32_t t)
{
	uint32_t h = fx_lookup(sl, t);
	if (h != FX_EMPTY && (sl.hr[h] & FX_ON_PATH) && sl.hp[h] == s) return true;
	h = fx_lookup(sl, s ^ 1);
	return h != FX_EMPTY && (sl.hr[h] & FX_ON_PATH) && sl.hp[h] == (t ^ 1);
}

// asg_bub_backtrack (asg.c:338-357) as NET deletions: everything visited goes, except what the best path revives.
// Returns false if the backtrack would revive a bit that was already deleted when v0 looked (never on a symmetric
// graph without multi-arcs, which is what asg_pop_bubble works on): the pass would not be deletion-only.
template <class V> FX_HD_NOINL bool fx_bub_backtrack(const V &g, uint32_t v0, const FxSlot &sl, const FxBubRes &w)
{
	bool mono = true;
	const uint32_t t = v0 + 1;
	uint32_t v = w.sink;
	do { // mark the best path sink -> ... -> child of v0
		const uint32_t h = fx_lookup(sl, v), u = sl.hp[h];
		sl.hr[h] |= FX_ON_PATH;
		if (g.st(v >> 1) <= v0) mono = false;
		{ // revived arcs u->v and v'->u' must have been live
			const uint64_t iu = g.idx[u]; const uint32_t off = (uint32_t)(iu >> 32);
			for (uint32_t i = 0; i < (uint32_t)iu; ++i) if (g.arc[off + i].v == v && g.at(off + i) <= v0) mono = false;
			const uint64_t ic = g.idx[v ^ 1]; const uint32_t offc = (uint32_t)(ic >> 32);
			for (uint32_t i = 0; i < (uint32_t)ic; ++i) if (g.arc[offc + i].v == (u ^ 1) && g.at(offc + i) <= v0) mono = false;
		}
		v = u;
	} while (v != v0);
	for (uint32_t i = 0; i < w.nb; ++i) { // reads: deleted unless one of their strands is on the best path
		const uint32_t x = sl.b[i];
		if (sl.hr[sl.bslot[i]] & FX_ON_PATH) continue;
		const uint32_t h = fx_lookup(sl, x ^ 1);
		if (h != FX_EMPTY && (sl.hr[h] & FX_ON_PATH)) continue;
		g.stamp_s(x >> 1, t);
	}
	for (uint32_t i = 0; i < w.ne; ++i) { // arcs and their complements
		const DArc a = g.arc[sl.e[i]];
		const uint32_t u = (uint32_t)(a.ul >> 32), x = a.v;
		if (!fx_revived(sl, u, x)) g.stamp_a(sl.e[i], t);
		if (!fx_revived(sl, x ^ 1, u ^ 1)) fx_arc_del(g, x ^ 1, u ^ 1, t);
	}
	return mono;
}

// one source of asg_pop_bubble's loop (asg.c:420-426): returns 1 = popped, 0 = nothing, -1 = scratch too small
template <class V> FX_HD int fx_bub_act(const V &g, uint32_t v0, uint32_t max_dist, const FxSlot &sl, uint32_t *n_tip, bool *mono)
{
	if (!fx_is_source(g, v0)) return 0;
	FxBubRes w;
	const int r = fx_bub_walk(g, v0, max_dist, sl, &w);
	if (r == 1) {
		if (!fx_bub_backtrack(g, v0, sl, w)) *mono = false;
		*n_tip = w.nT;
	}
	fx_bub_reset(sl, w.nb);
	return r;
}
