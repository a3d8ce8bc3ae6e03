/* gfa.c -- text writers and unitig sequence filling: the I/O-bound tail of the path, host C as in the
 * reference (asm.c:41-55 ma_sg_print, asm.c:64-116 ma_ug_destroy/ma_ug_print, asm.c:216-290 ma_ug_seq).
 * The output format is the parity contract: byte-identical S/L/a/x lines. */
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <unistd.h>
#include <pthread.h>
#include <sched.h>
#include "miniasm_b200.h"

void ma_ug_destroy(ma_ug_t *ug)
{
	size_t i;
	if (ug == 0) return;
	for (i = 0; i < ug->u.n; ++i) free(ug->u.a[i].a), free(ug->u.a[i].s);
	free(ug->u.a);
	asg_destroy(ug->g);
	free(ug);
}

void ma_sg_print(const asg_t *g, const sdict_t *d, const ma_sub_t *sub, FILE *fp)
{
	uint32_t i;
	for (i = 0; i < g->n_arc; ++i) {
		const asg_arc_t *a = &g->arc[i];
		uint32_t u = (uint32_t)(a->ul >> 32), v = a->v;
		const char *un = d->seq[u >> 1].name, *vn = d->seq[v >> 1].name;
		if (sub)
			fprintf(fp, "L\t%s:%d-%d\t%c\t%s:%d-%d\t%c\t%d:\tL1:i:%d\n", un, sub[u >> 1].s + 1, sub[u >> 1].e, "+-"[u & 1],
					vn, sub[v >> 1].s + 1, sub[v >> 1].e, "+-"[v & 1], a->ol, (uint32_t)a->ul);
		else
			fprintf(fp, "L\t%s\t%c\t%s\t%c\t%d:\tL1:i:%d\n", un, "+-"[u & 1], vn, "+-"[v & 1], a->ol, (uint32_t)a->ul);
	}
}

/* A small append buffer with hand-rolled integer formatting: the `a` lines are one per read, i.e. millions of
 * lines on real inputs, and fprintf dominated the host tail of the end-to-end path.  Output bytes are those of
 * the reference's fprintf formats (asm.c:77-116): "%d" of the 32-bit value, "utg%.6d" zero-padded to 6 digits. */
typedef struct { char *s; size_t l, m; FILE *fp; } obuf_t;

static inline void ob_room(obuf_t *b, size_t k)
{
	if (b->l + k <= b->m) return;
	if (b->fp == 0) { /* memory-only buffer (one per writer thread): grow */
		b->m = b->l + k > 2 * b->m ? b->l + k + (1 << 20) : 2 * b->m;
		b->s = (char*)realloc(b->s, b->m);
		return;
	}
	if (b->l) fwrite(b->s, 1, b->l, b->fp), b->l = 0;
	if (k > b->m) { b->m = k + (1 << 20); b->s = (char*)realloc(b->s, b->m); }
}
static inline void ob_c(obuf_t *b, char c) { b->s[b->l++] = c; }
static inline void ob_str(obuf_t *b, const char *s) { size_t n = strlen(s); ob_room(b, n + 64); memcpy(b->s + b->l, s, n); b->l += n; }
static inline void ob_int(obuf_t *b, int32_t v) /* "%d" */
{
	char t[12]; int n = 0;
	uint32_t x = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
	if (v < 0) ob_c(b, '-');
	do t[n++] = '0' + x % 10, x /= 10; while (x);
	while (n) ob_c(b, t[--n]);
}
static inline void ob_utg(obuf_t *b, uint32_t i, int circ) /* "utg%.6d%c" of i+1 */
{
	char t[12]; int n = 0;
	uint32_t x = i + 1;
	do t[n++] = '0' + x % 10, x /= 10; while (x);
	ob_c(b, 'u'); ob_c(b, 't'); ob_c(b, 'g');
	for (x = n; x < 6; ++x) ob_c(b, '0');
	while (n) ob_c(b, t[--n]);
	ob_c(b, "lc"[!!circ]);
}
static inline void ob_read(obuf_t *b, const sdict_t *d, const ma_sub_t *sub, uint32_t r) /* name or name:s+1-e */
{
	ob_str(b, d->seq[r].name);
	if (sub) { ob_c(b, ':'); ob_int(b, (int32_t)(sub[r].s + 1)); ob_c(b, '-'); ob_int(b, (int32_t)sub[r].e); }
}

/* S line, circularising L lines (item 0 of a unitig) and `a` lines (items 1..n) of unitig i, items [it0, it1).
 * off0 = layout offset of the first `a` line emitted. */
static void ug_emit_items(obuf_t *b, const ma_ug_t *ug, const sdict_t *d, const ma_sub_t *sub, uint32_t i, uint32_t it0, uint32_t it1, uint32_t off0)
{
	const ma_utg_t *p = &ug->u.a[i];
	uint32_t j, off = off0;
	if (it0 == 0 && it1 > 0) {
		ob_room(b, 256);
		ob_c(b, 'S'); ob_c(b, '\t'); ob_utg(b, i, p->circ); ob_c(b, '\t');
		if (p->s) { size_t n = strlen(p->s); ob_room(b, n + 256); memcpy(b->s + b->l, p->s, n); b->l += n; }
		else ob_c(b, '*');
		memcpy(b->s + b->l, "\tLN:i:", 6), b->l += 6; ob_int(b, (int32_t)p->len); ob_c(b, '\n');
		if (p->circ) {
			int k;
			for (k = 0; k < 2; ++k) {
				ob_c(b, 'L'); ob_c(b, '\t'); ob_utg(b, i, 1); ob_c(b, '\t'); ob_c(b, "+-"[k]); ob_c(b, '\t');
				ob_utg(b, i, 1); ob_c(b, '\t'); ob_c(b, "+-"[k]); ob_c(b, '\t'); ob_c(b, '0'); ob_c(b, 'M'); ob_c(b, '\n');
			}
		}
	}
	for (j = it0 ? it0 - 1 : 0; j + 1 < it1; ++j) {
		const uint32_t r = (uint32_t)(p->a[j] >> 33), l = (uint32_t)p->a[j];
		ob_room(b, 256);
		ob_c(b, 'a'); ob_c(b, '\t'); ob_utg(b, i, p->circ); ob_c(b, '\t'); ob_int(b, (int32_t)off); ob_c(b, '\t');
		ob_read(b, d, sub, r);
		ob_c(b, '\t'); ob_c(b, "+-"[p->a[j] >> 32 & 1]); ob_c(b, '\t'); ob_int(b, (int32_t)l); ob_c(b, '\n');
		off += l;
	}
}

/* The `a` lines are one per read and each touches three scattered host records (name, interval, layout entry),
 * so on million-read layouts the writer is bound by cache misses and integer formatting.  Large outputs are
 * therefore formatted by worker threads in blocks of UGW_BLOCK items (block b belongs to worker b mod T, two
 * reusable buffers per worker) while the calling thread writes the finished blocks to `fp` in order. */
#define UGW_BLOCK 4096
#define UGW_MAX_THREADS 16

typedef struct { uint32_t i, it, off; } ugw_cur_t; /* unitig, item inside it, layout offset of its next `a` line */

typedef struct {
	const ma_ug_t *ug; const sdict_t *d; const ma_sub_t *sub;
	const ugw_cur_t *cur;  /* cursor at the start of each block */
	uint64_t n_items;
	uint32_t n_blk, n_written;
	int n_thr, go;
	uint8_t *ready;
	obuf_t *buf;           /* 2 per worker */
	pthread_mutex_t mtx;
	pthread_cond_t cv;
} ugw_shared_t;

typedef struct { ugw_shared_t *sh; int k; } ugw_arg_t;

/* Blocks take ~0.1 ms to format, far less than a futex sleep/wake round trip, so the hand-over between the workers
 * and the writing thread polls (release/acquire flags), yielding the core once the wait gets long. */
static inline void ugw_relax(unsigned *spins)
{
	if (++*spins < 2000) {
#if defined(__x86_64__) || defined(__i386__)
		__builtin_ia32_pause();
#endif
	} else sched_yield();
}

static void ug_emit_block(obuf_t *b, const ugw_shared_t *sh, uint32_t blk)
{
	uint64_t left = sh->n_items - (uint64_t)blk * UGW_BLOCK;
	uint32_t i = sh->cur[blk].i, it = sh->cur[blk].it, off = sh->cur[blk].off;
	if (left > UGW_BLOCK) left = UGW_BLOCK;
	while (left) {
		const uint32_t n_it = sh->ug->u.a[i].n + 1;
		const uint32_t end = (uint64_t)(n_it - it) <= left ? n_it : it + (uint32_t)left;
		ug_emit_items(b, sh->ug, sh->d, sh->sub, i, it, end, off);
		left -= end - it;
		++i, it = 0, off = 0;
	}
}

static void *ug_worker(void *arg)
{
	ugw_shared_t *sh = ((ugw_arg_t*)arg)->sh;
	const int k = ((ugw_arg_t*)arg)->k;
	uint32_t b;
	int T;
	pthread_mutex_lock(&sh->mtx);
	while (!sh->go) pthread_cond_wait(&sh->cv, &sh->mtx);
	T = sh->n_thr;
	pthread_mutex_unlock(&sh->mtx);
	if (k >= T) return 0;
	for (b = k; b < sh->n_blk; b += T) {
		obuf_t *ob = &sh->buf[2 * k + (b / T & 1)];
		if (b >= 2u * T) { /* the buffer last held block b - 2T: wait until that one has been written out */
			unsigned spins = 0;
			while (__atomic_load_n(&sh->n_written, __ATOMIC_ACQUIRE) <= b - 2u * T) ugw_relax(&spins);
		}
		ob->l = 0;
		ug_emit_block(ob, sh, b);
		__atomic_store_n(&sh->ready[b], 1, __ATOMIC_RELEASE);
	}
	return 0;
}

static int ug_writer_threads(uint64_t n_items)
{
	long nc = sysconf(_SC_NPROCESSORS_ONLN);
	const char *e = getenv("MAB_WRITER_THREADS");
	int t = (int)(n_items / (8 * UGW_BLOCK)); /* at least 8 blocks per worker, else not worth starting threads */
	if (nc < 1) nc = 1;
	if (t > nc - 1) t = (int)nc - 1; /* the caller's thread does the writing */
	if (e && atoi(e) >= 0) t = atoi(e);
	if (t > UGW_MAX_THREADS) t = UGW_MAX_THREADS;
	return t;
}

void ma_ug_print(const ma_ug_t *ug, const sdict_t *d, const ma_sub_t *sub, FILE *fp)
{
	uint32_t i, j;
	uint64_t n_items = 0;
	int n_thr;
	obuf_t b = {0, 0, 0, fp};
	ob_room(&b, 1 << 20);
	for (i = 0; i < ug->u.n; ++i) n_items += (uint64_t)ug->u.a[i].n + 1;
	n_thr = ug_writer_threads(n_items);
	if (n_thr >= 1 && n_items > UGW_BLOCK && n_items / UGW_BLOCK < (1u << 31)) { /* segments, circularising links, read layout */
		ugw_shared_t sh;
		ugw_arg_t *arg = (ugw_arg_t*)calloc(n_thr, sizeof(ugw_arg_t));
		pthread_t *tid = (pthread_t*)calloc(n_thr, sizeof(pthread_t));
		ugw_cur_t *cur;
		uint32_t ui = 0, it = 0, off = 0, blk;
		int k, n_started = 0;
		memset(&sh, 0, sizeof(sh));
		sh.ug = ug, sh.d = d, sh.sub = sub, sh.n_items = n_items;
		sh.n_blk = (uint32_t)((n_items + UGW_BLOCK - 1) / UGW_BLOCK);
		cur = (ugw_cur_t*)malloc((size_t)sh.n_blk * sizeof(ugw_cur_t));
		for (blk = 0; blk < sh.n_blk; ++blk) { /* one serial pass over the layout: where each block starts */
			uint64_t left = n_items - (uint64_t)blk * UGW_BLOCK;
			if (left > UGW_BLOCK) left = UGW_BLOCK;
			cur[blk].i = ui, cur[blk].it = it, cur[blk].off = off;
			while (left) {
				const ma_utg_t *p = &ug->u.a[ui];
				const uint32_t n_it = p->n + 1;
				const uint32_t end = (uint64_t)(n_it - it) <= left ? n_it : it + (uint32_t)left;
				for (j = it ? it - 1 : 0; j + 1 < end; ++j) off += (uint32_t)p->a[j];
				left -= end - it;
				if (end == n_it) ++ui, it = 0, off = 0;
				else it = end;
			}
		}
		sh.cur = cur;
		sh.ready = (uint8_t*)calloc(sh.n_blk, 1);
		sh.buf = (obuf_t*)calloc(2 * (size_t)n_thr, sizeof(obuf_t));
		pthread_mutex_init(&sh.mtx, 0);
		pthread_cond_init(&sh.cv, 0);
		for (k = 0; k < n_thr; ++k) {
			arg[k].sh = &sh, arg[k].k = k;
			if (pthread_create(&tid[k], 0, ug_worker, &arg[k]) != 0) break;
			++n_started;
		}
		pthread_mutex_lock(&sh.mtx);
		sh.n_thr = n_started, sh.go = 1;
		pthread_cond_broadcast(&sh.cv);
		pthread_mutex_unlock(&sh.mtx);
		if (n_started == 0) { /* no thread could be started: format here */
			for (i = 0; i < ug->u.n; ++i) ug_emit_items(&b, ug, d, sub, i, 0, ug->u.a[i].n + 1, 0);
		} else {
			for (blk = 0; blk < sh.n_blk; ++blk) {
				const obuf_t *ob = &sh.buf[2 * (blk % n_started) + (blk / n_started & 1)];
				unsigned spins = 0;
				while (!__atomic_load_n(&sh.ready[blk], __ATOMIC_ACQUIRE)) ugw_relax(&spins);
				if (ob->l) fwrite(ob->s, 1, ob->l, fp);
				__atomic_store_n(&sh.n_written, blk + 1, __ATOMIC_RELEASE);
			}
		}
		for (k = 0; k < n_started; ++k) pthread_join(tid[k], 0);
		for (k = 0; k < 2 * n_thr; ++k) free(sh.buf[k].s);
		pthread_mutex_destroy(&sh.mtx);
		pthread_cond_destroy(&sh.cv);
		free(sh.buf); free(sh.ready); free(cur); free(arg); free(tid);
	} else {
		for (i = 0; i < ug->u.n; ++i) ug_emit_items(&b, ug, d, sub, i, 0, ug->u.a[i].n + 1, 0);
	}
	for (i = 0; i < ug->g->n_arc; ++i) { /* links between unitigs */
		const asg_arc_t *a = &ug->g->arc[i];
		const uint32_t u = (uint32_t)(a->ul >> 32), v = a->v;
		ob_room(&b, 256);
		ob_c(&b, 'L'); ob_c(&b, '\t'); ob_utg(&b, u >> 1, ug->u.a[u >> 1].circ); ob_c(&b, '\t'); ob_c(&b, "+-"[u & 1]); ob_c(&b, '\t');
		ob_utg(&b, v >> 1, ug->u.a[v >> 1].circ); ob_c(&b, '\t'); ob_c(&b, "+-"[v & 1]); ob_c(&b, '\t');
		ob_int(&b, (int32_t)a->ol); ob_c(&b, 'M'); memcpy(b.s + b.l, "\tSD:i:", 6), b.l += 6; ob_int(&b, (int32_t)(uint32_t)a->ul); ob_c(&b, '\n');
	}
	for (i = 0; i < ug->u.n; ++i) { /* per-unitig summary */
		const ma_utg_t *p = &ug->u.a[i];
		ob_room(&b, 512);
		ob_c(&b, 'x'); ob_c(&b, '\t');
		if (p->start == UINT32_MAX) {
			ob_utg(&b, i, 1); ob_c(&b, '\t'); ob_int(&b, (int32_t)p->len); ob_c(&b, '\t'); ob_int(&b, (int32_t)p->n); ob_c(&b, '\n');
		} else {
			ob_utg(&b, i, 0); ob_c(&b, '\t'); ob_int(&b, (int32_t)p->len); ob_c(&b, '\t'); ob_int(&b, (int32_t)p->n); ob_c(&b, '\t');
			ob_int(&b, (int32_t)(uint32_t)ug->g->idx[i << 1 | 1]); ob_c(&b, '\t'); ob_int(&b, (int32_t)(uint32_t)ug->g->idx[i << 1 | 0]); ob_c(&b, '\t');
			ob_read(&b, d, sub, p->start >> 1); ob_c(&b, '\t'); ob_c(&b, "+-"[p->start & 1]); ob_c(&b, '\t');
			ob_read(&b, d, sub, p->end >> 1); ob_c(&b, '\t'); ob_c(&b, "+-"[p->end & 1]); ob_c(&b, '\n');
		}
	}
	if (b.l) fwrite(b.s, 1, b.l, fp);
	free(b.s);
}

/* ---------------------------------------------------------------------------------------------
 * FASTA/FASTQ streaming (gzip or plain), semantics of kseq.h:160-234 as used by asm.c:236-290:
 * a record starts at '>' or '@'; the name ends at the first white space; sequence lines run until a
 * line starting with '>', '@' or '+'; after '+' quality bytes are consumed until they match the
 * sequence length.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
	gzFile fp;
	unsigned char *buf;
	int beg, end, eof, last; /* last: pending header character of the next record */
	kstring_t name, seq;
} fx_t;

#define FX_CHUNK 0x10000

static int fx_getc(fx_t *f)
{
	if (f->beg >= f->end) {
		if (f->eof) return -1;
		f->beg = 0;
		f->end = gzread(f->fp, f->buf, FX_CHUNK);
		if (f->end < FX_CHUNK) f->eof = 1;
		if (f->end <= 0) { f->end = 0; return -1; }
	}
	return f->buf[f->beg++];
}

static void ks_putc(kstring_t *s, int c)
{
	if (s->l + 2 > s->m) { s->m = s->m ? s->m << 1 : 256; s->s = (char*)realloc(s->s, s->m); }
	s->s[s->l++] = (char)c, s->s[s->l] = 0;
}

/* returns sequence length, or -1 at end of input */
static long fx_read(fx_t *f)
{
	int c;
	if (f->last == 0) {
		while ((c = fx_getc(f)) != -1 && c != '>' && c != '@');
		if (c == -1) return -1;
		f->last = c;
	}
	f->name.l = f->seq.l = 0;
	if (f->name.s) f->name.s[0] = 0;
	while ((c = fx_getc(f)) != -1 && !isspace(c)) ks_putc(&f->name, c);
	if (c == -1 && f->name.l == 0) return -1;
	if (c != -1 && c != '\n') while ((c = fx_getc(f)) != -1 && c != '\n'); /* comment */
	if (f->seq.s == 0) ks_putc(&f->seq, 0), f->seq.l = 0, f->seq.s[0] = 0;
	while ((c = fx_getc(f)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		ks_putc(&f->seq, c);
		while ((c = fx_getc(f)) != -1 && c != '\n') ks_putc(&f->seq, c); /* rest of the line */
		if (f->seq.l && f->seq.s[f->seq.l - 1] == '\r') f->seq.s[--f->seq.l] = 0;
	}
	f->last = (c == '>' || c == '@') ? c : 0;
	if (c == '+') {
		size_t ql = 0;
		while ((c = fx_getc(f)) != -1 && c != '\n');         /* rest of the '+' line */
		while (ql < f->seq.l && (c = fx_getc(f)) != -1) {     /* quality lines */
			if (c == '\n' || c == '\r') continue;
			++ql;
		}
		f->last = 0;
	}
	return (long)f->seq.l;
}

typedef struct { uint32_t utg:31, ori:1, start, len; } utg_slot_t;

int ma_ug_seq(ma_ug_t *g, const sdict_t *d, const ma_sub_t *sub, const char *fn)
{
	static unsigned char comp[256];
	static const char pairs[] = "ATCGBVDHKMRY";
	fx_t f;
	utg_slot_t *slot;
	uint32_t i, j;
	gzFile fp = fn && strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(fileno(stdin), "r");
	if (fp == 0) return -1;
	for (i = 0; i < 256; ++i) comp[i] = i < 128 ? (unsigned char)i : 'N';
	for (i = 0; pairs[i]; i += 2) {
		comp[(int)pairs[i]] = pairs[i + 1], comp[(int)pairs[i + 1]] = pairs[i];
		comp[tolower(pairs[i])] = tolower(pairs[i + 1]), comp[tolower(pairs[i + 1])] = tolower(pairs[i]);
	}
	comp['U'] = 'A', comp['u'] = 'a', comp[96] = 64; /* table quirks of asm.c:224-233 */
	memset(&f, 0, sizeof(f));
	f.fp = fp, f.buf = (unsigned char*)malloc(FX_CHUNK);

	slot = (utg_slot_t*)calloc(d->n_seq ? d->n_seq : 1, sizeof(utg_slot_t));
	for (i = 0; i < g->u.n; ++i) {
		ma_utg_t *u = &g->u.a[i];
		uint32_t off = 0;
		u->s = (char*)calloc(1, (size_t)u->len + 1);
		memset(u->s, 'N', u->len);
		for (j = 0; j < u->n; ++j) {
			utg_slot_t *t = &slot[u->a[j] >> 33];
			if (t->len != 0) { /* asm.c:246 asserts it: a read lies on one unitig, once */
				fprintf(stderr, "[E::%s] read %u is placed twice in the layout\n", __func__, (uint32_t)(u->a[j] >> 33));
				abort();
			}
			t->utg = i, t->ori = u->a[j] >> 32 & 1, t->start = off, t->len = (uint32_t)u->a[j];
			off += t->len;
		}
	}
	while (fx_read(&f) >= 0) {
		int32_t id = sd_get(d, f.name.s ? f.name.s : "");
		const utg_slot_t *t;
		char *dst;
		const char *src = f.seq.s;
		size_t sl = f.seq.l;
		if (id < 0 || slot[id].len == 0) continue;
		t = &slot[id];
		dst = g->u.a[t->utg].s + t->start;
		/* asm.c:263 asserts that the record covers the kept interval; a reads file that does not belong to the PAF must not
		 * make us read outside the record (without `sub` the reference would: we refuse there too) */
		if (sub ? sub[id].e > f.seq.l : t->len > f.seq.l) {
			fprintf(stderr, "[E::%s] sequence '%s' has %ld bases, the layout needs %u: wrong reads file?\n", __func__, f.name.s, (long)f.seq.l,
					sub ? sub[id].e : t->len);
			abort();
		}
		if (sub) src += sub[id].s, sl = sub[id].e - sub[id].s;
		if (!t->ori) memcpy(dst, src, t->len);
		else for (i = 0; i < t->len; ++i) dst[i] = (char)comp[(unsigned char)src[sl - 1 - i]];
	}
	free(slot);
	free(f.buf); free(f.name.s); free(f.seq.s);
	gzclose(fp);
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * -R prefilter: one streaming pass that names the reads clearly contained in a much longer read
 * (hit.c:38-68).  Dictionary work on a text stream: host C, like the reference.
 * --------------------------------------------------------------------------------------------- */
sdict_t *ma_hit_no_cont(const char *fn, int min_span, int min_match, int max_hang, float int_frac)
{
	paf_file_t *fp = paf_open(fn);
	paf_rec_t r;
	sdict_t *d;
	if (fp == 0) {
		fprintf(stderr, "[E::%s] could not open PAF file %s\n", __func__, fn);
		exit(1);
	}
	memset(&r, 0, sizeof(r));
	d = sd_init();
	while (paf_read(fp, &r) >= 0) {
		int l5, l3;
		if (r.qe - r.qs < (uint32_t)min_span || r.te - r.ts < (uint32_t)min_span || (int)r.ml < min_match) continue;
		l5 = r.rev ? r.tl - r.te : r.ts;
		l3 = r.rev ? r.ts : r.tl - r.te;
		if (r.ql >> 1 > r.tl) { /* query at least twice as long: is the target inside it? */
			if (l5 > max_hang >> 2 || l3 > max_hang >> 2 || r.te - r.ts < r.tl * int_frac) continue;
			if ((int)r.qs - l5 > max_hang << 1 && (int)(r.ql - r.qe) - l3 > max_hang << 1) sd_put(d, r.tn, r.tl);
		} else if (r.ql < r.tl >> 1) {
			if (r.qs > (uint32_t)(max_hang >> 2) || r.ql - r.qe > (uint32_t)(max_hang >> 2) || r.qe - r.qs < r.ql * int_frac) continue;
			if (l5 - (int)r.qs > max_hang << 1 && l3 - (int)(r.ql - r.qe) > max_hang << 1) sd_put(d, r.qn, r.ql);
		}
	}
	paf_close(fp);
	if (ma_verbose >= 3) fprintf(stderr, "[M::%s::%s] dropped %d contained reads\n", __func__, sys_timestamp(), d->n_seq);
	return d;
}
