/* sdict.c -- read-name dictionary: name -> dense id in order of first appearance
 * (reference: sdict.c:8-86; ids define the vertex numbering).
 * The index is a private open-addressing table (FNV-1a, linear probing) stored behind sdict_t::h. */
#include <stdlib.h>
#include <string.h>
#include "miniasm_b200.h"

typedef struct {
	uint32_t n_slot, n_used;
	int32_t *slot;           /* NULL = index not built yet (bulk-created dictionary): built on the first sd_get/sd_put */
	char *block; size_t block_size; /* names of a bulk-created dictionary live in one allocation */
} sd_index_t;

static uint64_t sd_hash_str(const char *s)
{
	uint64_t h = 1469598103934665603ULL;
	for (; *s; ++s) h = (h ^ (uint8_t)*s) * 1099511628211ULL;
	return h ^ h >> 29;
}

static sd_index_t *sx_new(uint32_t n_slot)
{
	sd_index_t *x = (sd_index_t*)calloc(1, sizeof(sd_index_t));
	x->n_slot = n_slot;
	x->slot = (int32_t*)malloc((size_t)n_slot * 4);
	memset(x->slot, 0xff, (size_t)n_slot * 4);
	return x;
}

static void sx_free(sd_index_t *x) { if (x) { free(x->slot); free(x->block); free(x); } }

static int sx_owns(const sd_index_t *x, const char *name) { return x && x->block && name >= x->block && name < x->block + x->block_size; }

static void sx_insert_id(sd_index_t *x, const sdict_t *d, int32_t id)
{
	uint32_t m = x->n_slot - 1, k = (uint32_t)sd_hash_str(d->seq[id].name) & m;
	while (x->slot[k] >= 0) k = (k + 1) & m;
	x->slot[k] = id, ++x->n_used;
}

static void sx_grow(sdict_t *d)
{
	sd_index_t *old = (sd_index_t*)d->h, *x = sx_new(old->n_slot << 1);
	uint32_t i;
	for (i = 0; i < old->n_slot; ++i)
		if (old->slot[i] >= 0) sx_insert_id(x, d, old->slot[i]);
	x->block = old->block, x->block_size = old->block_size, old->block = 0;
	sx_free(old);
	d->h = x;
}

static void sx_materialise(sdict_t *d) /* build the deferred index of a bulk-created dictionary */
{
	sd_index_t *x = (sd_index_t*)d->h;
	uint32_t i, n_slot = 1024;
	if (x == 0 || x->slot) return;
	while ((uint64_t)n_slot * 7 < (uint64_t)d->n_seq * 10 + 10) n_slot <<= 1;
	x->n_slot = n_slot, x->n_used = 0;
	x->slot = (int32_t*)malloc((size_t)n_slot * 4);
	memset(x->slot, 0xff, (size_t)n_slot * 4);
	for (i = 0; i < d->n_seq; ++i) sx_insert_id(x, d, (int32_t)i);
}

/* Bulk constructor used by the fused level: n names packed back to back (NUL terminated) in `block`, which the
 * dictionary takes over; ids are the positions.  The hash index is only built if somebody looks a name up. */
sdict_t *sd_from_packed(char *block, size_t block_size, uint32_t n, const uint32_t *len)
{
	sdict_t *d = (sdict_t*)calloc(1, sizeof(sdict_t));
	sd_index_t *x = (sd_index_t*)calloc(1, sizeof(sd_index_t));
	char *p = block;
	uint32_t i;
	d->n_seq = d->m_seq = n;
	d->seq = (sd_seq_t*)malloc((size_t)(n ? n : 1) * sizeof(sd_seq_t));
	for (i = 0; i < n; ++i) {
		d->seq[i].name = p, d->seq[i].len = len[i], d->seq[i].aux = 0, d->seq[i].del = 0;
		p += strlen(p) + 1;
	}
	x->block = block, x->block_size = block_size;
	d->h = x;
	return d;
}

sdict_t *sd_init(void)
{
	sdict_t *d = (sdict_t*)calloc(1, sizeof(sdict_t));
	d->h = sx_new(1024);
	return d;
}

void sd_destroy(sdict_t *d)
{
	uint32_t i;
	if (d == 0) return;
	for (i = 0; i < d->n_seq; ++i)
		if (!sx_owns((sd_index_t*)d->h, d->seq[i].name)) free(d->seq[i].name);
	sx_free((sd_index_t*)d->h);
	free(d->seq);
	free(d);
}

int32_t sd_get(const sdict_t *d, const char *name)
{
	const sd_index_t *x = (const sd_index_t*)d->h;
	uint32_t m, k;
	if (x == 0) return -1;
	if (x->slot == 0) sx_materialise((sdict_t*)d);
	m = x->n_slot - 1;
	for (k = (uint32_t)sd_hash_str(name) & m; x->slot[k] >= 0; k = (k + 1) & m)
		if (strcmp(d->seq[x->slot[k]].name, name) == 0) return x->slot[k];
	return -1;
}

int32_t sd_put(sdict_t *d, const char *name, uint32_t len)
{
	sd_index_t *x;
	sd_seq_t *s;
	int32_t id = sd_get(d, name);
	if (id >= 0) return id; /* the length of a known name is not re-checked (sdict.c:43) */
	if (d->n_seq == d->m_seq) {
		d->m_seq = d->m_seq ? d->m_seq << 1 : 16;
		d->seq = (sd_seq_t*)realloc(d->seq, (size_t)d->m_seq * sizeof(sd_seq_t));
	}
	s = &d->seq[d->n_seq];
	s->name = strdup(name), s->len = len, s->aux = 0, s->del = 0;
	x = (sd_index_t*)d->h;
	if (x == 0) d->h = x = sx_new(1024);
	if ((uint64_t)(x->n_used + 1) * 10 > (uint64_t)x->n_slot * 7) sx_grow(d), x = (sd_index_t*)d->h;
	sx_insert_id(x, d, (int32_t)d->n_seq);
	return (int32_t)d->n_seq++;
}

/* drop deleted entries, keep the order, rebuild the index; returns old id -> new id (-1 if dropped) */
int32_t *sd_squeeze(sdict_t *d)
{
	int32_t *map = (int32_t*)calloc(d->n_seq ? d->n_seq : 1, 4);
	uint32_t i, j, n_slot = 1024;
	sd_index_t *x, *old = (sd_index_t*)d->h;
	for (i = j = 0; i < d->n_seq; ++i) {
		if (d->seq[i].del) { if (!sx_owns(old, d->seq[i].name)) free(d->seq[i].name); map[i] = -1; }
		else d->seq[j] = d->seq[i], map[i] = (int32_t)j++;
	}
	d->n_seq = j;
	while ((uint64_t)n_slot * 7 < (uint64_t)j * 10 + 10) n_slot <<= 1;
	x = sx_new(n_slot);
	if (old) x->block = old->block, x->block_size = old->block_size, old->block = 0;
	sx_free(old);
	d->h = x;
	for (i = 0; i < j; ++i) sx_insert_id(x, d, (int32_t)i);
	return map;
}
