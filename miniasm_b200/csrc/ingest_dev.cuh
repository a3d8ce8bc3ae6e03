// ingest_dev.cuh -- PAF text -> sorted hit array, entirely on the GPU (reference: paf.c:34-67
// paf_parse/paf_read, hit.c:70-107 ma_hit_read, sdict.c:27-45 sd_put).
#pragma once
#include "mab_common.cuh"
#include "hit_dev.cuh"

// name table of the reads, in id order (id = rank of first appearance among filter-passing lines,
// query before target: hit.c:87-90 + sdict.c:27-45)
struct DNames {
	uint32_t n_seq = 0;
	uint64_t *off = nullptr;     // byte offset of the name in the PAF text (first occurrence)
	uint32_t *nlen = nullptr;    // name length in bytes
	uint32_t *slen = nullptr;    // sequence length recorded at first appearance
};

struct IngestStats { uint64_t n_lines, n_parsed, n_hits, n_seq, tot_len; uint32_t max_qs_bits; int hash_retries; };

// d_text: the PAF bytes in device memory.  On return `h` holds the sorted hits (ma_hit_sort order, stable) and
// `names` the dictionary.  Exits loudly on malformed sizes (> 2^31 hits per GPU).
void ingest_paf(MabDev &d, const char *d_text, size_t len, int min_span, int min_match, int bi_dir,
                DHits &h, DNames &names, IngestStats &st);
void names_free(MabDev &d, DNames &n);

struct ShardComm;
// Sharded variant: this rank's byte range of the PAF in, the hits of the reads this rank owns out.
// *name_text_out receives a device buffer with all read names packed (names.off indexes it); the caller frees it.
void ingest_paf_sharded(MabDev &d, ShardComm &sc, const char *d_text, size_t len, int min_span, int min_match, int bi_dir,
                        DHits &h, DNames &names, char **name_text_out, IngestStats &st);
