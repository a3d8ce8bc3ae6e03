/* pafgen -- deterministic synthetic all-vs-all PAF generator (layout model of SURVEY.md section 8d).
 *
 * Reads are intervals on a random linear genome; every pair of reads whose genomic intersection is
 * at least `min_olap` bp yields one PAF line (12 columns, mapq 255) with a random query/target role.
 * Everything is driven by a splitmix64 stream seeded from `seed`, so the same options give the same
 * bytes on every machine: tests pin golden outputs of the reference to (options, sha256 of the PAF).
 *
 * Build:  gcc -O2 -o pafgen pafgen.c            (CLI)
 *         gcc -O2 -fPIC -shared -DPAFGEN_LIB -o libpafgen.so pafgen.c   (ctypes: pafgen_generate)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>

typedef struct {
	uint32_t n_reads;        /* number of reads (before hot-spot extras) */
	uint32_t len_min, len_max; /* read length ~ U[len_min, len_max] */
	double   coverage;       /* genome length = n_reads * mean_len / coverage */
	uint32_t min_olap;       /* minimum genomic intersection to emit a line */
	uint32_t jitter;         /* each alignment end is pulled in by U[0, jitter] bp, per read */
	uint64_t seed;
	uint32_t n_hot, hot_reads, hot_span; /* skew: n_hot loci, hot_reads extra reads each, starts within hot_span */
	uint32_t dup_ppm;        /* per-million chance to emit a line twice (multi-arcs) */
	uint32_t self_ppm;       /* per-million chance per read to emit a self hit */
	uint32_t internal_ppm;   /* per-million chance to truncate one alignment end by up to 3 kb (internal match) */
	uint32_t lowid_ppm;      /* per-million chance of a low-identity line (ml = bl/50) */
	uint32_t shuffle;        /* shuffle line order */
	uint32_t flip_ppm;       /* per-million chance to swap q/t role is fixed at 50%; this adds CR line ends */
	uint32_t name_base;      /* added to every read number: disjoint names for the partitions of a multi-GPU run */
	uint32_t part, n_parts;  /* n_parts > 1: emit only part `part` of the PAF -- the lines whose first read (in genomic order) has index
	                          * in [n*part/n_parts, n*(part+1)/n_parts).  The parts, concatenated in order, are the bytes of the whole
	                          * PAF: the random stream does not depend on what is emitted.  Not with `shuffle`. */
} pafgen_opt_t;

typedef struct { uint64_t n_lines, n_bytes, genome_len; uint32_t n_reads_total; } pafgen_stat_t;

static inline uint64_t sm64(uint64_t *s)
{
	uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
static inline uint64_t rnd_below(uint64_t *s, uint64_t n) { return n ? sm64(s) % n : 0; }

typedef struct { uint64_t start; uint32_t len, name; uint8_t rev; } rd_t;

static int cmp_rd(const void *a, const void *b)
{
	const rd_t *x = (const rd_t*)a, *y = (const rd_t*)b;
	if (x->start != y->start) return x->start < y->start ? -1 : 1;
	return x->name < y->name ? -1 : x->name > y->name;
}

typedef struct { char *s; size_t l, m; } sbuf_t;

static inline void sb_need(sbuf_t *b, size_t k)
{
	if (b->l + k > b->m) {
		b->m = (b->l + k) * 3 / 2 + 4096;
		b->s = (char*)realloc(b->s, b->m);
		if (!b->s) { fprintf(stderr, "pafgen: out of memory\n"); exit(1); }
	}
}
static inline void sb_u32(sbuf_t *b, uint32_t x)
{
	char t[12]; int n = 0;
	do { t[n++] = '0' + x % 10; x /= 10; } while (x);
	while (n) b->s[b->l++] = t[--n];
}
static inline void sb_c(sbuf_t *b, char c) { b->s[b->l++] = c; }

/* local coordinates of genomic [a,b) on a read placed at [s,s+len) with strand rev */
static inline void to_local(const rd_t *r, uint64_t a, uint64_t b, uint32_t *ls, uint32_t *le)
{
	if (!r->rev) *ls = (uint32_t)(a - r->start), *le = (uint32_t)(b - r->start);
	else *ls = (uint32_t)(r->start + r->len - b), *le = (uint32_t)(r->start + r->len - a);
}

static void emit(sbuf_t *b, const rd_t *q, const rd_t *t, uint32_t qs, uint32_t qe, uint32_t ts, uint32_t te, int lowid, int cr)
{
	uint32_t qspan = qe - qs, tspan = te - ts, bl = qspan > tspan ? qspan : tspan;
	uint32_t ml = lowid ? bl / 50 : bl / 5;
	sb_need(b, 160);
	sb_c(b, 'r'); sb_u32(b, q->name); sb_c(b, '\t'); sb_u32(b, q->len); sb_c(b, '\t'); sb_u32(b, qs); sb_c(b, '\t'); sb_u32(b, qe);
	sb_c(b, '\t'); sb_c(b, q->rev == t->rev ? '+' : '-'); sb_c(b, '\t');
	sb_c(b, 'r'); sb_u32(b, t->name); sb_c(b, '\t'); sb_u32(b, t->len); sb_c(b, '\t'); sb_u32(b, ts); sb_c(b, '\t'); sb_u32(b, te);
	sb_c(b, '\t'); sb_u32(b, ml); sb_c(b, '\t'); sb_u32(b, bl); sb_c(b, '\t'); sb_c(b, '2'); sb_c(b, '5'); sb_c(b, '5');
	if (cr) sb_c(b, '\r');
	sb_c(b, '\n');
}

/* Generates the PAF text into a malloc'd buffer (*out, caller frees with pafgen_free). */
size_t pafgen_generate(const pafgen_opt_t *o, char **out, pafgen_stat_t *st)
{
	uint64_t rng = o->seed * 0x2545F4914F6CDD1DULL + 0x1234567ULL, G;
	uint32_t n = o->n_reads + o->n_hot * o->hot_reads, i, j, k, part_lo, part_hi;
	double mean_len = 0.5 * ((double)o->len_min + o->len_max);
	rd_t *r = (rd_t*)malloc((size_t)n * sizeof(rd_t));
	uint32_t *perm = (uint32_t*)malloc((size_t)n * 4);
	sbuf_t b = {0, 0, 0};
	uint64_t n_lines = 0, *line_off = 0;
	size_t line_m = 0;

	G = (uint64_t)((double)o->n_reads * mean_len / (o->coverage > 0 ? o->coverage : 30.0));
	if (G < (uint64_t)o->len_max * 2) G = (uint64_t)o->len_max * 2;
	for (i = 0; i < n; ++i) perm[i] = i;
	for (i = n; i > 1; --i) { j = (uint32_t)rnd_below(&rng, i); k = perm[i-1]; perm[i-1] = perm[j]; perm[j] = k; }
	for (i = 0; i < o->n_reads; ++i) {
		r[i].len = o->len_min + (uint32_t)rnd_below(&rng, (uint64_t)o->len_max - o->len_min + 1);
		r[i].start = rnd_below(&rng, G - r[i].len + 1);
		r[i].rev = sm64(&rng) >> 63;
		r[i].name = perm[i] + o->name_base;
	}
	for (k = 0; k < o->n_hot; ++k) {
		uint64_t locus = rnd_below(&rng, G - o->len_max - o->hot_span);
		for (j = 0; j < o->hot_reads; ++j, ++i) {
			r[i].len = o->len_min + (uint32_t)rnd_below(&rng, (uint64_t)o->len_max - o->len_min + 1);
			r[i].start = locus + rnd_below(&rng, o->hot_span + 1);
			r[i].rev = sm64(&rng) >> 63;
			r[i].name = perm[i] + o->name_base;
		}
	}
	free(perm);
	qsort(r, n, sizeof(rd_t), cmp_rd);

	if (o->n_parts > 1) part_lo = (uint32_t)((uint64_t)n * o->part / o->n_parts), part_hi = (uint32_t)((uint64_t)n * (o->part + 1) / o->n_parts);
	else part_lo = 0, part_hi = n;
	for (i = 0; i < n; ++i) {
		uint64_t ei = r[i].start + r[i].len;
		const int mine = i >= part_lo && i < part_hi;
		if (i >= part_hi && !o->shuffle) break; /* nothing of the later reads belongs to this part */
		if (o->self_ppm && rnd_below(&rng, 1000000) < o->self_ppm) { /* self hit: a palindromic-looking or shifted one */
			uint32_t w = r[i].len / 2;
			const int pal = sm64(&rng) & 1;
			if (mine) {
				if (o->shuffle) { if (n_lines == line_m) { line_m = line_m ? line_m * 2 : 1024; line_off = (uint64_t*)realloc(line_off, line_m * 8); } line_off[n_lines] = b.l; }
				if (pal) { rd_t t = r[i]; t.rev = !t.rev; emit(&b, &r[i], &t, 100, 100 + w, 100, 100 + w, 0, 0); }
				else emit(&b, &r[i], &r[i], 0, w, r[i].len - w, r[i].len, 0, 0);
				++n_lines;
			}
		}
		for (j = i + 1; j < n && r[j].start + o->min_olap <= ei; ++j) {
			uint64_t a = r[j].start, e = ei < r[j].start + r[j].len ? ei : r[j].start + r[j].len;
			uint64_t a1 = a, e1 = e, a2 = a, e2 = e;
			const rd_t *q, *t;
			uint32_t qs, qe, ts, te, dup, lowid = 0;
			if (e - a < o->min_olap) continue;
			if (o->jitter) {
				a1 += rnd_below(&rng, o->jitter + 1); e1 -= rnd_below(&rng, o->jitter + 1);
				a2 += rnd_below(&rng, o->jitter + 1); e2 -= rnd_below(&rng, o->jitter + 1);
			}
			if (o->internal_ppm && rnd_below(&rng, 1000000) < o->internal_ppm) {
				uint64_t cut = 1 + rnd_below(&rng, 3000);
				if (sm64(&rng) & 1) a1 += cut, a2 += cut; else e1 -= cut, e2 -= cut;
			}
			if (o->lowid_ppm && rnd_below(&rng, 1000000) < o->lowid_ppm) lowid = 1;
			if (a1 + 50 >= e1 || a2 + 50 >= e2) continue;
			if (sm64(&rng) & 1) q = &r[i], t = &r[j]; else q = &r[j], t = &r[i];
			if (q == &r[i]) to_local(q, a1, e1, &qs, &qe), to_local(t, a2, e2, &ts, &te);
			else to_local(q, a2, e2, &qs, &qe), to_local(t, a1, e1, &ts, &te);
			dup = (o->dup_ppm && rnd_below(&rng, 1000000) < o->dup_ppm) ? 2 : 1;
			while (dup--) {
				const int cr = o->flip_ppm && rnd_below(&rng, 1000000) < o->flip_ppm;
				if (!mine) continue;
				if (o->shuffle) { if (n_lines == line_m) { line_m = line_m ? line_m * 2 : 1024; line_off = (uint64_t*)realloc(line_off, line_m * 8); } line_off[n_lines] = b.l; }
				emit(&b, q, t, qs, qe, ts, te, lowid, cr);
				++n_lines;
			}
		}
	}
	if (o->shuffle && n_lines > 1) {
		sbuf_t c = {0, 0, 0};
		uint64_t *ord = (uint64_t*)malloc(n_lines * 8), x, t;
		sb_need(&c, b.l + 1);
		for (x = 0; x < n_lines; ++x) ord[x] = x;
		for (x = n_lines; x > 1; --x) { uint64_t y = rnd_below(&rng, x); t = ord[x-1]; ord[x-1] = ord[y]; ord[y] = t; }
		for (x = 0; x < n_lines; ++x) {
			uint64_t beg = line_off[ord[x]], end = ord[x] + 1 < n_lines ? line_off[ord[x] + 1] : b.l;
			memcpy(c.s + c.l, b.s + beg, end - beg); c.l += end - beg;
		}
		free(ord); free(b.s); b = c;
	}
	free(line_off);
	free(r);
	if (st) st->n_lines = n_lines, st->n_bytes = b.l, st->genome_len = G, st->n_reads_total = n;
	*out = b.s;
	return b.l;
}

void pafgen_free(char *p) { free(p); }

void pafgen_defaults(pafgen_opt_t *o)
{
	memset(o, 0, sizeof(*o));
	o->n_reads = 10000; o->len_min = o->len_max = 10000; o->coverage = 62.5; o->min_olap = 2000; o->seed = 1; o->hot_span = 8000;
}

#ifndef PAFGEN_LIB
int main(int argc, char *argv[])
{
	pafgen_opt_t o;
	pafgen_stat_t st;
	char *buf;
	size_t len;
	int c;
	pafgen_defaults(&o);
	while ((c = getopt(argc, argv, "n:l:L:c:m:j:s:H:R:W:d:S:I:D:xC:B:P:")) >= 0) {
		if (c == 'n') o.n_reads = strtoul(optarg, 0, 10);
		else if (c == 'l') o.len_min = atoi(optarg);
		else if (c == 'L') o.len_max = atoi(optarg);
		else if (c == 'c') o.coverage = atof(optarg);
		else if (c == 'm') o.min_olap = atoi(optarg);
		else if (c == 'j') o.jitter = atoi(optarg);
		else if (c == 's') o.seed = strtoull(optarg, 0, 10);
		else if (c == 'H') o.n_hot = atoi(optarg);
		else if (c == 'R') o.hot_reads = atoi(optarg);
		else if (c == 'W') o.hot_span = atoi(optarg);
		else if (c == 'd') o.dup_ppm = atoi(optarg);
		else if (c == 'S') o.self_ppm = atoi(optarg);
		else if (c == 'I') o.internal_ppm = atoi(optarg);
		else if (c == 'D') o.lowid_ppm = atoi(optarg);
		else if (c == 'C') o.flip_ppm = atoi(optarg);
		else if (c == 'x') o.shuffle = 1;
		else if (c == 'B') o.name_base = strtoul(optarg, 0, 10);
		else if (c == 'P') { char *e; o.part = strtoul(optarg, &e, 10); o.n_parts = *e == '/' ? strtoul(e + 1, 0, 10) : 1; }
	}
	if (o.len_max < o.len_min) o.len_max = o.len_min;
	if (o.n_parts > 1 && (o.shuffle || o.part >= o.n_parts)) { fprintf(stderr, "pafgen: -P part/parts needs part < parts and no -x\n"); return 1; }
	len = pafgen_generate(&o, &buf, &st);
	fwrite(buf, 1, len, stdout);
	fprintf(stderr, "[pafgen] reads=%u genome=%lu lines=%lu bytes=%lu\n", st.n_reads_total, (unsigned long)st.genome_len,
			(unsigned long)st.n_lines, (unsigned long)st.n_bytes);
	pafgen_free(buf);
	return 0;
}
#endif
