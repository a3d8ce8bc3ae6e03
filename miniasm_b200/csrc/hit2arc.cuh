// hit2arc.cuh -- the hit classifier shared by ma_hit_flt, ma_hit_contained and ma_sg_gen.
//
// Restates ma_hit2arc (miniasm.h:86-104) with the C integer-conversion rules of the original written out:
// the reference mixes int32, uint32 and 31-bit fields, so which comparisons are signed and which are
// unsigned is part of the behaviour (hard part 2).  One float32 product decides the
// "internal match" test; it is computed with an explicit round-to-nearest multiply (no FMA contraction).
#pragma once
#include "mab_common.cuh"

#define MAB_HT_INT        (-1)
#define MAB_HT_QCONT      (-2)
#define MAB_HT_TCONT      (-3)
#define MAB_HT_SHORT_OVLP (-4)

// ql, tl: lengths of the (trimmed) query / target reads.  On a non-negative return *arc holds the arc.
__host__ __device__ __forceinline__ int mab_hit2arc(const DHit &h, int ql, int tl, int max_hang, float int_frac, int min_ovlp, DArc *arc)
{
	const int32_t qs = (int32_t)(uint32_t)h.qns;
	const uint32_t qe = h.qe, rev = h.ml_rev >> 31;
	int32_t tl5, tl3; // target overhang beyond the query's 5' / 3' end, on the query strand
	if (rev) tl5 = (int32_t)((uint32_t)tl - h.te), tl3 = (int32_t)h.ts;
	else tl5 = (int32_t)h.ts, tl3 = (int32_t)((uint32_t)tl - h.te);
	const uint32_t q3 = (uint32_t)ql - qe;                       // query bases right of the hit (unsigned in the reference)
	const int32_t ext5 = qs < tl5 ? qs : tl5;                    // signed compare
	const int32_t ext3 = (int32_t)(q3 < (uint32_t)tl3 ? q3 : (uint32_t)tl3); // unsigned compare
	const uint32_t span = qe - (uint32_t)qs;
	const uint32_t full = span + (uint32_t)ext5 + (uint32_t)ext3;
#ifdef __CUDA_ARCH__
	const float bound = __fmul_rn((float)full, int_frac);
#else
	const float bound = (float)full * int_frac;
#endif
	if (ext5 > max_hang || ext3 > max_hang || (float)span < bound) return MAB_HT_INT;
	if (qs <= tl5 && q3 <= (uint32_t)tl3) return MAB_HT_QCONT;    // query contained (tested first: ties delete both reads)
	if (qs >= tl5 && q3 >= (uint32_t)tl3) return MAB_HT_TCONT;    // target contained
	uint32_t u, v, l;
	if (qs > tl5) u = 0, v = rev, l = (uint32_t)qs - (uint32_t)tl5;
	else u = 1, v = !rev, l = q3 - (uint32_t)tl3;
	if (full < (uint32_t)min_ovlp || h.te - h.ts + (uint32_t)ext5 + (uint32_t)ext3 < (uint32_t)min_ovlp) return MAB_HT_SHORT_OVLP;
	u |= (uint32_t)(h.qns >> 32) << 1, v |= h.tn << 1;
	arc->ul = (uint64_t)u << 32 | l;
	arc->v = v;
	arc->ol_del = ((uint32_t)ql - l) & 0x7fffffffu;
	return (int)l;
}
