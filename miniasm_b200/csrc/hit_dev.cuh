// hit_dev.cuh -- stage (i): per-read coverage trimming and containment filtering of PAF hits on the GPU.
// Counterpart of hit.c:109-256 and the shared classifier ma_hit2arc (miniasm.h:86-104).
#pragma once
#include "mab_common.cuh"
#include "asg_dev.cuh"
#include <functional>

// sharded runs: sums the per-rank counters of a log line over all ranks (installed by mab_select_sharded; null = single GPU)
typedef void (*MabCountHook)(void *ctx, unsigned long long *v, int n);
extern thread_local MabCountHook mab_count_hook;
extern thread_local void *mab_count_hook_ctx;

struct DHits {
	DHit *a = nullptr, *a2 = nullptr;   // hit array + ping-pong buffer for compaction
	size_t n = 0, m = 0;
	uint32_t n_seq = 0;
};

struct HitArcParams { int max_hang; float int_frac; int min_ovlp; };

void dh_reserve(MabDev &d, DHits &h, size_t m);
void dh_free(MabDev &d, DHits &h);

// ma_hit_sort (hit.c:19-22): sort by the 64-bit qns (query id, then query start); stable.
void dh_sort(MabDev &d, DHits &h, uint32_t max_len_bits);

// ma_hit_sub (hit.c:109-160).  sub_out: n_seq entries, fully written (zeros for reads heading no group).
// Returns the number of reads that keep an interval ("query sequences remain after sub").
uint64_t dh_sub(MabDev &d, const DHits &h, int min_dp, float min_iden, int end_clip, DSub *sub_out);

// ma_hit_cut (hit.c:162-193): clip hits to the kept intervals, drop short ones; returns the new count.
size_t dh_cut(MabDev &d, DHits &h, const DSub *reg, int min_span);

// ma_hit_flt (hit.c:195-216): drop internal / short hits; cov as the reference computes it (logged only).
size_t dh_flt(MabDev &d, DHits &h, const DSub *sub, int max_hang, int min_ovlp, float *cov);

// ma_sub_merge (hit.c:218-223)
void dh_sub_merge(MabDev &d, uint32_t n_sub, DSub *a, const DSub *b);

// ma_hit_contained (hit.c:225-256) + ma_hit_mark_unused (hit.c:24-36) + the id part of sd_squeeze
// (sdict.c:69-86).  seq_del: per-read deletion flags of the dictionary on entry (may be null = none).
// On return sub is compacted in place, hits renumbered/compacted, map_out[old] = new id or -1, and
// h.n_seq is the surviving read count.  Returns the new hit count.
// With cut_reg != null the call first applies ma_hit_cut(cut_reg, min_span) to the hits (fused sweep, one compaction).
// `exchange` (sharded runs) is called between the flag pass and the renumbering with (sub, used, n_seq).
size_t dh_contained(MabDev &d, DHits &h, DSub *sub, const uint8_t *seq_del, const HitArcParams &p, int32_t *map_out,
                    const DSub *cut_reg = nullptr, int min_span = 0,
                    const std::function<void(DSub*, uint8_t*, uint32_t)> *exchange = nullptr);
// ma_sg_gen without the final asg_cleanup: seq table + sorted local arcs (sharded runs clean up after exchanging seq flags)
void dh_sg_emit(MabDev &d, const DHits &h, const uint32_t *len, const uint8_t *del, const HitArcParams &p, DGraph &g);
// ma_hit_cut immediately followed by ma_hit_flt on the same table (main.c:123-125), fused
size_t dh_cut_flt(MabDev &d, DHits &h, const DSub *sub, int min_span, int max_hang, int min_ovlp, float *cov);

// ma_sg_gen (asm.c:9-39): lens/del per read -> graph with arcs emitted in hit order, then asg_cleanup.
void dh_sg_gen(MabDev &d, const DHits &h, const uint32_t *len, const uint8_t *del, const HitArcParams &p, DGraph &g);
