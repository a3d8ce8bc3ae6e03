/* paf.c -- line reader and field parser for PAF text, gzip or plain, "-" = stdin
 * (reference: paf.c:9-67 on top of kseq.h:101-149; format PAF.md:7-28).
 *
 * Behaviour kept: lines end at '\n'; one trailing '\r' is dropped when the line is longer than one
 * byte; fields split on TAB only; columns 2-4 and 7-11 go through strtol(.,10) and are truncated into
 * uint32 (ml into 31 bits); rev = first byte of column 5 is '-'; a line with fewer than 10 fields is
 * skipped silently; a line with exactly 10 fields keeps the previous record's bl. */
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "miniasm_b200.h"

#define PAF_CHUNK 0x10000

typedef struct {
	gzFile fp;
	unsigned char *buf;
	int beg, end, eof;
} paf_stream_t;

paf_file_t *paf_open(const char *fn)
{
	paf_file_t *pf;
	paf_stream_t *st;
	gzFile fp = fn && strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(fileno(stdin), "r");
	if (fp == 0) return 0;
	st = (paf_stream_t*)calloc(1, sizeof(paf_stream_t));
	st->fp = fp;
	st->buf = (unsigned char*)malloc(PAF_CHUNK);
	pf = (paf_file_t*)calloc(1, sizeof(paf_file_t));
	pf->fp = st;
	return pf;
}

int paf_close(paf_file_t *pf)
{
	paf_stream_t *st;
	if (pf == 0) return 0;
	st = (paf_stream_t*)pf->fp;
	gzclose(st->fp);
	free(st->buf); free(st);
	free(pf->buf.s);
	free(pf);
	return 0;
}

/* next line into pf->buf (NUL terminated); returns its length or -1 at end of input */
static int paf_next_line(paf_file_t *pf)
{
	paf_stream_t *st = (paf_stream_t*)pf->fp;
	kstring_t *s = &pf->buf;
	s->l = 0;
	if (st->beg >= st->end && st->eof) return -1;
	for (;;) {
		int i;
		if (st->beg >= st->end) {
			if (st->eof) break;
			st->beg = 0;
			st->end = gzread(st->fp, st->buf, PAF_CHUNK);
			if (st->end < PAF_CHUNK) st->eof = 1;
			if (st->end <= 0) { st->end = 0; break; }
		}
		for (i = st->beg; i < st->end && st->buf[i] != '\n'; ++i);
		if (s->m - s->l < (size_t)(i - st->beg + 1)) {
			s->m = s->l + (i - st->beg) + 1;
			s->m += s->m >> 1;
			s->s = (char*)realloc(s->s, s->m);
		}
		memcpy(s->s + s->l, st->buf + st->beg, i - st->beg);
		s->l += i - st->beg;
		st->beg = i + 1;
		if (i < st->end) break; /* hit the newline */
	}
	if (s->s == 0) s->m = 1, s->s = (char*)calloc(1, 1);
	else if (s->l > 1 && s->s[s->l - 1] == '\r') --s->l;
	s->s[s->l] = 0;
	return (int)s->l;
}

static int paf_fields(int l, char *s, paf_rec_t *r)
{
	int i, k = 0;
	char *f = s, *e;
	for (i = 0; i <= l; ++i) {
		if (i < l && s[i] != '\t') continue;
		s[i] = 0;
		switch (k) {
			case 0: r->qn = f; break;
			case 1: r->ql = strtol(f, &e, 10); break;
			case 2: r->qs = strtol(f, &e, 10); break;
			case 3: r->qe = strtol(f, &e, 10); break;
			case 4: r->rev = (*f == '-'); break;
			case 5: r->tn = f; break;
			case 6: r->tl = strtol(f, &e, 10); break;
			case 7: r->ts = strtol(f, &e, 10); break;
			case 8: r->te = strtol(f, &e, 10); break;
			case 9: r->ml = strtol(f, &e, 10); break;
			case 10: r->bl = strtol(f, &e, 10); break;
			default: break;
		}
		++k;
		f = i < l ? s + i + 1 : 0;
	}
	return k < 10 ? -1 : 0;
}

int paf_read(paf_file_t *pf, paf_rec_t *r)
{
	int len;
	while ((len = paf_next_line(pf)) >= 0)
		if (paf_fields(len, pf->buf.s, r) == 0) return 0;
	return len;
}
