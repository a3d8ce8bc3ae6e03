// gfa_dev.cuh -- ma_ug_print's text (asm.c:77-116) formatted on the GPU; see gfa_dev.cu.
#pragma once
#include "mab_common.cuh"
#include "clean_dev.cuh"

// orig: current read id -> original id (null = identity); noff/nlen/name_text: names by original id; sub: kept intervals
// by current id (null = names without the :s-e suffix).  *d_text_out: device buffer with the text (free with d.free).
size_t dg_gfa_text(MabDev &d, const DUnitigs &ug, const uint32_t *orig, const uint64_t *noff, const uint32_t *nlen, const char *name_text,
                   const DSub *sub, char **d_text_out);
