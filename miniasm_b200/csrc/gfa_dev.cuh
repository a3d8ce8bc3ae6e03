// gfa_dev.cuh -- ma_ug_print's text (asm.c:77-116) formatted on the GPU; see gfa_dev.cu.
#pragma once
#include "mab_common.cuh"
#include "clean_dev.cuh"

// orig: current read id -> original id (null = identity); noff/nlen/name_text: names by original id; sub: kept intervals
// by current id (null = names without the :s-e suffix).  *d_text_out: device buffer with the text (free with d.free).
// seq_pos (device, n_utg entries) non-null: every S line reserves `len` bytes, pre-filled with 'N', for the unitig sequence and
// seq_pos[i] receives their offset in the text (ugseq_dev.cu gathers the bases there); *ioff_out then keeps the exclusive sum of the
// item lengths (caller frees).
size_t dg_gfa_text(MabDev &d, const DUnitigs &ug, const uint32_t *orig, const uint64_t *noff, const uint32_t *nlen, const char *name_text,
                   const DSub *sub, char **d_text_out, uint64_t *seq_pos = nullptr, uint32_t **ioff_out = nullptr);
