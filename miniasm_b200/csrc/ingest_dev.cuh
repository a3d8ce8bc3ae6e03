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

struct IngestStats { uint64_t n_lines, n_parsed, n_hits, n_seq, tot_len, n_dropped; uint32_t max_qs_bits; int hash_retries; };
struct NoContParams { int max_hang; float int_frac; }; // -R (ma_hit_no_cont, hit.c:38-68)

// d_text: the PAF bytes in device memory.  On return `h` holds the sorted hits (ma_hit_sort order, stable) and
// `names` the dictionary.  Exits loudly on malformed sizes (> 2^31 hits per GPU).
// nocont != nullptr: the reference's -R, i.e. reads clearly contained in a much longer read are dropped together with every
// line that names them before ids are assigned (st.n_dropped = their number).
void ingest_paf(MabDev &d, const char *d_text, size_t len, int min_span, int min_match, int bi_dir,
                DHits &h, DNames &names, IngestStats &st, const NoContParams *nocont = nullptr);
// load + ingest overlapped: host_text -> d_text in chunks on MabDev::copy_stream while arrived chunks are scanned and parsed
void ingest_paf_stream(MabDev &d, char *d_text, const char *host_text, size_t len, int min_span, int min_match, int bi_dir,
                       DHits &h, DNames &names, IngestStats &st);
void names_free(MabDev &d, DNames &n);
// byte offsets of the line starts of a text in device memory (len > 0); free with d.free.  start[n_lines] is not set.
uint64_t *dev_line_starts(MabDev &d, const char *d_text, size_t len, uint64_t *n_lines_out);
// FNV-1a + finaliser over the bytes [s, t) of p, never 0: the read-name hash of the dictionaries
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
	return k;
}

__host__ __device__ __forceinline__ uint64_t name_hash(const char *p, uint32_t s, uint32_t t, uint64_t seed)
{
	uint64_t h = 1469598103934665603ULL ^ seed;
	for (uint32_t i = s; i < t; ++i) h = (h ^ (uint8_t)p[i]) * 1099511628211ULL;
	h = fmix64(h);
	return h ? h : 1;
}



struct ShardComm;
// Sharded variant: this rank's byte range of the PAF in, the hits of the reads this rank owns out.
// *name_text_out receives a device buffer with all read names packed (names.off indexes it); the caller frees it.
// host_text != nullptr: the bytes are still on the host and cross PCIe in chunks while the arrived chunks are parsed.
void ingest_paf_sharded(MabDev &d, ShardComm &sc, char *d_text, size_t len, int min_span, int min_match, int bi_dir,
                        DHits &h, DNames &names, char **name_text_out, IngestStats &st, const char *host_text = nullptr);
