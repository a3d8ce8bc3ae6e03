// ugseq_dev.cuh -- ma_ug_seq (asm.c:236-290) on the GPU; see ugseq_dev.cu.
#pragma once
#include "mab_common.cuh"
#include "clean_dev.cuh"

enum { UGSEQ_UNSUPPORTED = -2, UGSEQ_SHORT_RECORD = -3 };

// index of a FASTA/FASTQ text in device memory
struct DReadsIndex {
	uint64_t *start = nullptr;     // line starts
	uint8_t *type = nullptr;       // per line: empty / header / '+' / sequence
	uint32_t *slen = nullptr;      // per line: length without the line end
	uint64_t *cum = nullptr;       // FASTA-like only: sequence bytes before line i (n_lines + 1 entries)
	uint64_t *hdr_line = nullptr;  // per record: its header line
	uint64_t n_lines = 0, n_rec = 0;
	int fq4 = 0;
};

// 0, or UGSEQ_UNSUPPORTED when the text is not one of the two layouts the parallel parser proves (caller: host reader)
int dg_reads_index(MabDev &d, const char *text, size_t len, DReadsIndex &ix);
void dg_reads_free(MabDev &d, DReadsIndex &ix);
// Gathers the bases of every layout item into out + seq_pos[unitig] + (offset of the item in its unitig).  ioff = exclusive sum
// (mod 2^32) of the item lengths; names by ORIGINAL read id (orig: current -> original, null = identity); sub by current id.
// Returns 0, or UGSEQ_SHORT_RECORD if a record is shorter than the interval the layout keeps of it (asm.c:263 asserts).
int dg_ugseq_fill(MabDev &d, const char *text, size_t len, const DReadsIndex &ix, const DUnitigs &ug, const uint32_t *ioff,
                  uint32_t n_seq, const uint32_t *orig, const uint64_t *noff, const uint32_t *nlen, const char *ntext, const DSub *sub,
                  const uint64_t *seq_pos, char *out);
