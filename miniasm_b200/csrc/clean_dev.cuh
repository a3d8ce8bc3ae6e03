// clean_dev.cuh -- stage (iii): tip / internal / bi-loop cutting, bubble popping and unitig construction on
// the device-resident graph.  Counterpart of asg.c:199-433 and asm.c:121-210.
#pragma once
#include "asg_dev.cuh"

uint32_t dg_cut_tip(MabDev &d, DGraph &g, int max_ext);        // asg.c:238-254
uint32_t dg_cut_internal(MabDev &d, DGraph &g, int max_ext);   // asg.c:256-272
uint32_t dg_cut_biloop(MabDev &d, DGraph &g, int max_ext);     // asg.c:274-306
uint64_t dg_pop_bubble(MabDev &d, DGraph &g, int max_dist);    // asg.c:312-433 (pops | trimmed tips << 32)

// statistics of the speculative rounds (DESIGN.md "sequential passes"): rounds and candidates of the last call
struct CleanStats { uint32_t rounds, committed;   // sweeps / actions of the last pass
                    uint32_t passes, max_rounds, sum_rounds, sum_committed; }; // accumulated since mab_clean_totals_reset (one layout)
extern thread_local CleanStats g_clean_stats;

// Unitigs (asm.c:121-210).  Device result, flattened:
//   utg_meta[i] = {len, circ, start, end, n, first}   items[first .. first+n) = vertex<<32 | length
struct DUtgMeta { uint32_t len, circ, start, end, n, first; };
struct DUnitigs {
	uint32_t n_utg = 0;
	uint64_t n_items = 0;
	DUtgMeta *meta = nullptr;
	uint64_t *items = nullptr;
	DGraph g;                      // unitig graph (asm.c:181-207), cleaned up
};
void dg_ug_gen(MabDev &d, const DGraph &g, DUnitigs &ug);
void dg_ug_free(MabDev &d, DUnitigs &ug);
