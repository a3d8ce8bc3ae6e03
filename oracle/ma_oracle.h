/* ma_oracle.h -- ABI structs of the reference seam for the CPU restatement (ORACLE, test infrastructure only).
 * Layouts follow sdict.h:6-15, asg.h:7-23, miniasm.h:12-55 of lh3/miniasm v0.3-r179. */
#ifndef MA_ORACLE_H
#define MA_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <sys/types.h>

typedef struct { char *name; uint32_t len, aux:31, del:1; } sd_seq_t;
typedef struct { uint32_t n_seq, m_seq; sd_seq_t *seq; void *h; } sdict_t;
typedef struct { uint64_t ul; uint32_t v; uint32_t ol:31, del:1; } asg_arc_t;
typedef struct { uint32_t len:31, del:1; } asg_seq_t;
typedef struct { uint32_t m_arc, n_arc:31, is_srt:1; asg_arc_t *arc; uint32_t m_seq, n_seq:31, is_symm:1; asg_seq_t *seq; uint64_t *idx; } asg_t;
typedef struct {
	int min_span, min_match, min_dp; float min_iden; int max_hang, min_ovlp; float int_frac;
	int gap_fuzz, n_rounds, bub_dist, max_ext; float min_ovlp_drop_ratio, max_ovlp_drop_ratio, final_ovlp_drop_ratio;
} ma_opt_t;
typedef struct { uint64_t qns; uint32_t qe, tn, ts, te; uint32_t ml:31, rev:1; uint32_t bl:31, del:1; } ma_hit_t;
typedef struct { uint32_t s:31, del:1, e; } ma_sub_t;
typedef struct { uint32_t len:31, circ:1; uint32_t start, end; uint32_t m, n; uint64_t *a; char *s; } ma_utg_t;
typedef struct { size_t n, m; ma_utg_t *a; } ma_utg_v;
typedef struct { ma_utg_v u; asg_t *g; } ma_ug_t;

void ma_opt_init(ma_opt_t *o);
sdict_t *sd_init(void);
void sd_destroy(sdict_t *d);
int32_t sd_get(const sdict_t *d, const char *name);
int32_t sd_put(sdict_t *d, const char *name, uint32_t len);
int32_t *sd_squeeze(sdict_t *d);
sdict_t *ma_hit_no_cont(const char *fn, int min_span, int min_match, int max_hang, float int_frac);
ma_hit_t *ma_hit_read(const char *fn, int min_span, int min_match, sdict_t *d, size_t *n, int bi_dir, const sdict_t *excl);
ma_sub_t *ma_hit_sub(int min_dp, float min_iden, int end_clip, size_t n, const ma_hit_t *a, size_t n_sub);
size_t ma_hit_cut(const ma_sub_t *reg, int min_span, size_t n, ma_hit_t *a);
size_t ma_hit_flt(const ma_sub_t *sub, int max_hang, int min_ovlp, size_t n, ma_hit_t *a, float *cov);
void ma_sub_merge(size_t n_sub, ma_sub_t *a, const ma_sub_t *b);
size_t ma_hit_contained(const ma_opt_t *opt, sdict_t *d, ma_sub_t *sub, size_t n, ma_hit_t *a);
asg_t *ma_sg_gen(const ma_opt_t *opt, const sdict_t *d, const ma_sub_t *sub, size_t n_hits, const ma_hit_t *hit);
asg_t *asg_init(void);
void asg_destroy(asg_t *g);
void asg_seq_set(asg_t *g, int sid, int len, int del);
void asg_arc_rm(asg_t *g);
void asg_arc_sort(asg_t *g);
void asg_arc_index(asg_t *g);
void asg_cleanup(asg_t *g);
void asg_symm(asg_t *g);
int asg_arc_del_multi(asg_t *g);
int asg_arc_del_asymm(asg_t *g);
int asg_arc_del_trans(asg_t *g, int fuzz);
int asg_arc_del_short(asg_t *g, float ratio);
int asg_cut_tip(asg_t *g, int max_ext);
int asg_cut_internal(asg_t *g, int max_ext);
int asg_cut_biloop(asg_t *g, int max_ext);
int asg_pop_bubble(asg_t *g, int max_dist);
ma_ug_t *ma_ug_gen(asg_t *g);
int ma_ug_seq(ma_ug_t *ug, const sdict_t *d, const ma_sub_t *sub, const char *fn);
void ma_ug_destroy(ma_ug_t *ug);
void ma_sg_print(const asg_t *g, const sdict_t *d, const ma_sub_t *sub, FILE *fp);
void ma_ug_print(const ma_ug_t *ug, const sdict_t *d, const ma_sub_t *sub, FILE *fp);
#endif
