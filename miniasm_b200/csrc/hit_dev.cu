// hit_dev.cu -- stage (i) on the GPU: ma_hit_sort / ma_hit_sub / ma_hit_cut / ma_hit_flt / ma_sub_merge /
// ma_hit_contained / ma_sg_gen (hit.c:19-36,109-256; asm.c:9-39).  Hits are 32-byte records moved with
// two 128-bit accesses; the per-read interval tables (8 B/read) are gathered through L2.
#include "hit_dev.cuh"
#include "hit2arc.cuh"
#include <cub/cub.cuh>
#include <functional>

extern "C" int ma_verbose;              // hit.c prints its [M::fn::timestamp] lines only when ma_verbose >= 3
extern "C" const char *sys_timestamp(void);
#define ma_verbose_dev (mab_mute ? 0 : ma_verbose)

// Sharded runs: the counts in the [M::...] lines are per rank; the caller installs a hook that sums them over the ranks (a collective:
// every rank reaches every print site, muted or not).  Unset = single GPU, nothing to do.
thread_local MabCountHook mab_count_hook = nullptr;
thread_local void *mab_count_hook_ctx = nullptr;
static inline void sum_ranks(unsigned long long *v, int n) { if (mab_count_hook) mab_count_hook(mab_count_hook_ctx, v, n); }

__device__ __forceinline__ DHit ld_hit(const DHit *p)
{
	const uint4 *q = reinterpret_cast<const uint4*>(p);
	uint4 a = __ldg(q), b = __ldg(q + 1);
	DHit h;
	h.qns = (uint64_t)a.y << 32 | a.x; h.qe = a.z; h.tn = a.w;
	h.ts = b.x; h.te = b.y; h.ml_rev = b.z; h.bl_del = b.w;
	return h;
}
__device__ __forceinline__ DHit ld_hit_rw(const DHit *p) // for kernels that also store to the array
{
	const uint4 *q = reinterpret_cast<const uint4*>(p);
	uint4 a = q[0], b = q[1];
	DHit h;
	h.qns = (uint64_t)a.y << 32 | a.x; h.qe = a.z; h.tn = a.w;
	h.ts = b.x; h.te = b.y; h.ml_rev = b.z; h.bl_del = b.w;
	return h;
}
__device__ __forceinline__ void st_hit(DHit *p, const DHit &h)
{
	uint4 *q = reinterpret_cast<uint4*>(p);
	q[0] = make_uint4((uint32_t)h.qns, (uint32_t)(h.qns >> 32), h.qe, h.tn);
	q[1] = make_uint4(h.ts, h.te, h.ml_rev, h.bl_del);
}

static inline uint32_t bits_for(uint64_t x) { uint32_t b = 0; while (x) ++b, x >>= 1; return b ? b : 1; }

void dh_reserve(MabDev &d, DHits &h, size_t m)
{
	if (m <= h.m) return;
	DHit *na = mab_alloc<DHit>(d, m), *nb = mab_alloc<DHit>(d, m);
	if (h.n) MAB_CUDA(cudaMemcpyAsync(na, h.a, h.n * sizeof(DHit), cudaMemcpyDeviceToDevice, d.stream));
	d.free(h.a); d.free(h.a2);
	h.a = na, h.a2 = nb, h.m = m;
}

void dh_free(MabDev &d, DHits &h)
{
	d.free(h.a); d.free(h.a2);
	h = DHits();
}

// ---------------------------------------------------------------------------------------------
// Stable compaction of hits by a byte flag array.  32-byte records: the generic library selection moves them at
// ~2.2 TB/s (ncu: 2.8 ms per 100 M hits), so the hit path has its own three-step compaction -- per-tile keep counts
// (flags only), a scan of the ~n/2048 tile counts, and a scatter in which every warp row reads 1 KB of contiguous
// records and writes them behind the tile's offset.  Order of the kept hits is the input order.
// MAB_CUB_SELECT=1 switches back to cub::DeviceSelect::Flagged.
// ---------------------------------------------------------------------------------------------
constexpr int SEL_ROWS = 8, SEL_THREADS = 256, SEL_TILE = SEL_ROWS * SEL_THREADS;
static_assert(SEL_ROWS * (SEL_THREADS / 32) == 64, "k_sel_scatter scans exactly two counts per lane of one warp");

__global__ void __launch_bounds__(SEL_THREADS)
k_sel_count(const uint8_t *__restrict__ flag, size_t n, uint32_t *__restrict__ tile_cnt, unsigned long long *__restrict__ total)
{
	__shared__ unsigned s_w[SEL_THREADS / 32];
	const size_t base = (size_t)blockIdx.x * SEL_TILE;
	unsigned c = 0;
	#pragma unroll
	for (int k = 0; k < SEL_ROWS; ++k) {
		const size_t i = base + (size_t)k * SEL_THREADS + threadIdx.x;
		c += (i < n && flag[i] != 0) ? 1u : 0u;
	}
	c = __reduce_add_sync(0xffffffffu, c);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned t = 0;
		#pragma unroll
		for (int w = 0; w < SEL_THREADS / 32; ++w) t += s_w[w];
		tile_cnt[blockIdx.x] = t;
		if (t) atomicAdd(total, (unsigned long long)t);
	}
}

__global__ void __launch_bounds__(SEL_THREADS)
k_sel_scatter(const DHit *__restrict__ in, const uint8_t *__restrict__ flag, size_t n, const uint32_t *__restrict__ tile_off,
              DHit *__restrict__ out, unsigned long long *__restrict__ n_out)
{
	constexpr int NW = SEL_THREADS / 32;
	__shared__ uint32_t s_off[SEL_ROWS * NW];   // (row, warp) -> kept records of the tile before that warp row
	__shared__ uint32_t s_total;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const size_t base = (size_t)blockIdx.x * SEL_TILE;
	unsigned bal[SEL_ROWS];
	#pragma unroll
	for (int k = 0; k < SEL_ROWS; ++k) {
		const size_t i = base + (size_t)k * SEL_THREADS + tid;
		const bool f = i < n && flag[i] != 0;
		bal[k] = __ballot_sync(0xffffffffu, f);
		if (lane == 0) s_off[k * NW + warp] = __popc(bal[k]);
	}
	__syncthreads();
	if (warp == 0) { // exclusive scan of the SEL_ROWS * NW = 64 counts in record order, two per lane
		const uint32_t c0 = s_off[2 * lane], c1 = s_off[2 * lane + 1];
		uint32_t inc = c0 + c1;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		const uint32_t exc = inc - (c0 + c1);
		s_off[2 * lane] = exc, s_off[2 * lane + 1] = exc + c0;
		if (lane == 31) s_total = inc;
	}
	__syncthreads();
	const uint32_t tbase = tile_off[blockIdx.x];
	#pragma unroll
	for (int k = 0; k < SEL_ROWS; ++k) {
		if (bal[k] >> lane & 1u) {
			const size_t i = base + (size_t)k * SEL_THREADS + tid;
			const uint32_t pos = tbase + s_off[k * NW + warp] + __popc(bal[k] & ((1u << lane) - 1u));
			st_hit(out + pos, ld_hit(in + i));
		}
	}
	if (blockIdx.x == gridDim.x - 1 && tid == 0) *n_out = (unsigned long long)tbase + s_total;
}

static size_t select_hits(MabDev &d, DHits &h, const uint8_t *flag)
{
	if (h.n == 0) return 0;
	static const bool use_cub = getenv("MAB_CUB_SELECT") && atoi(getenv("MAB_CUB_SELECT")) != 0;
	unsigned long long *d_n = d.d_scal + SC_NSEL;
	if (use_cub || h.n >= (1ull << 32)) {
		size_t tb = 0;
		cub::DeviceSelect::Flagged(nullptr, tb, h.a, flag, h.a2, d_n, (int64_t)h.n, d.stream);
		void *tmp = d.tmp(tb);
		cub::DeviceSelect::Flagged(tmp, tb, h.a, flag, h.a2, d_n, (int64_t)h.n, d.stream);
		++d.n_lib;
	} else {
		const uint32_t n_tile = (uint32_t)((h.n + SEL_TILE - 1) / SEL_TILE);
		uint32_t *cnt = mab_alloc<uint32_t>(d, n_tile), *off = mab_alloc<uint32_t>(d, n_tile);
		d.zero_scal(SC_NSEL);
		MAB_LAUNCH(d, k_sel_count, n_tile, SEL_THREADS, 0, flag, h.n, cnt, d_n);
		if ((size_t)d.get_scal(SC_NSEL) == h.n) { // every record stays (clean data: the common case of ma_hit_cut/flt): nothing to move
			d.free(cnt); d.free(off);
			return h.n;
		}
		size_t tb = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, (int)n_tile, d.stream);
		void *tmp = d.tmp(tb);
		cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, (int)n_tile, d.stream);
		++d.n_lib;
		MAB_LAUNCH(d, k_sel_scatter, n_tile, SEL_THREADS, 0, h.a, flag, h.n, off, h.a2, d_n);
		d.free(cnt); d.free(off);   // stream-ordered arena: reuse is ordered behind the kernels above
	}
	size_t n = (size_t)d.get_scal(SC_NSEL);
	DHit *t = h.a; h.a = h.a2; h.a2 = t;
	h.n = n;
	return n;
}

// ---------------------------------------------------------------------------------------------
// ma_hit_sort: (key = qid << lb | qs, payload = position) radix sort, then a 32-byte gather
// ---------------------------------------------------------------------------------------------
__global__ void k_hit_keys(const DHit *a, size_t n, uint32_t lb, uint64_t *key, uint32_t *pos)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint64_t q = a[i].qns;
		key[i] = lb >= 32 ? q : ((q >> 32) << lb | (uint32_t)q);
		pos[i] = (uint32_t)i;
	}
}
__global__ void k_hit_gather(const DHit *a, const uint32_t *pos, size_t n, DHit *out)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		st_hit(out + i, ld_hit(a + pos[i]));
}

void dh_sort(MabDev &d, DHits &h, uint32_t max_len_bits)
{
	if (h.n < 2) return;
	if (h.n >= (1ull << 32)) { fprintf(stderr, "[E::miniasm_b200] more than 2^32 hits on one GPU\n"); exit(73); }
	uint32_t lb = max_len_bits > 32 ? 32 : max_len_bits;
	int end_bit = (int)(lb + bits_for(h.n_seq ? h.n_seq - 1 : 0));
	uint64_t *ka = mab_alloc<uint64_t>(d, h.n), *kb = mab_alloc<uint64_t>(d, h.n);
	uint32_t *pa = mab_alloc<uint32_t>(d, h.n), *pb = mab_alloc<uint32_t>(d, h.n);
	MAB_LAUNCH(d, k_hit_keys, mab_grid(h.n, 256), 256, 0, h.a, h.n, lb, ka, pa);
	cub::DoubleBuffer<uint64_t> dk(ka, kb);
	cub::DoubleBuffer<uint32_t> dp(pa, pb);
	size_t tb = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dp, (int64_t)h.n, 0, end_bit, d.stream);
	void *tmp = d.tmp(tb);
	cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dp, (int64_t)h.n, 0, end_bit, d.stream);
	++d.n_lib;
	MAB_LAUNCH(d, k_hit_gather, mab_grid(h.n, 256), 256, 0, h.a, dp.Current(), h.n, h.a2);
	DHit *t = h.a; h.a = h.a2; h.a2 = t;
	d.free(ka); d.free(kb); d.free(pa); d.free(pb);
}

// ---------------------------------------------------------------------------------------------
// ma_hit_sub (hit.c:109-160)
//   group = maximal run of equal query id;  per group collect (qs+clip)<<1 and (qe-clip)<<1|1 of the hits
//   with tn != qid and ml >= bl*min_iden (float32) and qe > qs;  sort;  sweep the depth;  keep the FIRST
//   longest stretch with depth >= min_dp;  sub = {start-clip, end+clip} or del.
// GPU shape: every hit emits two 64-bit keys  qid<<33 | invalid<<32 | endpoint  (invalid keys sink to the
// end of their group), one device-wide radix sort orders all groups at once, then one warp per group
// sweeps its 2*count keys with a warp scan.
// ---------------------------------------------------------------------------------------------
__global__ void k_group_bounds(const DHit *a, size_t n, uint32_t *g32)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint32_t q = (uint32_t)(a[i].qns >> 32);
		if (i == 0 || (uint32_t)(a[i - 1].qns >> 32) != q) g32[2 * (size_t)q + 1] = (uint32_t)i;
		if (i == n - 1 || (uint32_t)(a[i + 1].qns >> 32) != q) g32[2 * (size_t)q] = (uint32_t)(i + 1);
	}
}

__global__ void k_sub_keys(const DHit *a, size_t n, float min_iden, uint32_t clip, uint64_t *key)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit h = ld_hit(a + i);
		uint32_t qid = (uint32_t)(h.qns >> 32);
		int ml = (int)(h.ml_rev & 0x7fffffffu), bl = (int)(h.bl_del & 0x7fffffffu);
		bool skip = h.tn == qid || (float)ml < __fmul_rn((float)bl, min_iden);
		uint32_t qs = (uint32_t)h.qns + clip, qe = h.qe - clip;
		uint64_t base = (uint64_t)qid << 33;
		if (!skip && qe > qs) {
			key[2 * i] = base | (uint32_t)(qs << 1);
			key[2 * i + 1] = base | (uint32_t)(qe << 1 | 1);
		} else {
			key[2 * i] = key[2 * i + 1] = base | 1ull << 32;
		}
	}
}

__global__ void __launch_bounds__(256)
k_sub_sweep(const uint64_t *key, const uint64_t *grp, uint32_t n_seq, int min_dp, uint32_t clip, DSub *sub, unsigned long long *n_remained)
{
	const int lane = threadIdx.x & 31;
	const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	unsigned remained = 0;
	for (uint32_t r = wid; r < n_seq; r += nw) {
		const uint64_t g = grp[r];
		const uint32_t end = (uint32_t)g, first = (uint32_t)(g >> 32);
		if (end == 0) continue;                      // read heads no group: stays {0,0,del=0}
		const size_t k0 = 2 * (size_t)first;
		const uint32_t nk = 2 * (end - first);
		int depth = 0;
		uint32_t carry_start = 0;
		unsigned long long best = 0;                 // len<<32 | ~index  (max = longest, earliest on ties)
		uint32_t best_end = 0;
		for (uint32_t c = 0; c < nk; c += 32) {
			const uint32_t i = c + lane;
			uint64_t k = i < nk ? key[k0 + i] : (1ull << 32);
			const bool valid = !(k >> 32 & 1);
			const uint32_t pos = (uint32_t)k >> 1;
			int delta = valid ? ((k & 1) ? -1 : 1) : 0, dp = delta;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, dp, o); if (lane >= o) dp += t; }
			dp += depth;
			const int old = dp - delta;
			const bool up = valid && old < min_dp && dp >= min_dp;
			const bool down = valid && old >= min_dp && dp < min_dp;
			const unsigned upm = __ballot_sync(0xffffffffu, up);
			const unsigned below = upm & ((1u << lane) - 1);
			const int src = below ? 31 - __clz(below) : 0;
			uint32_t st = __shfl_sync(0xffffffffu, pos, src);
			if (!below) st = carry_start;
			unsigned long long cand = 0;
			if (down) cand = (unsigned long long)(pos - st) << 32 | (0xffffffffu - i);
			unsigned long long m = cand;
			#pragma unroll
			for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
			if (m > best && (m >> 32) != 0) {
				const unsigned who = __ballot_sync(0xffffffffu, down && cand == m);
				best = m;
				best_end = __shfl_sync(0xffffffffu, pos, __ffs(who) - 1);
			}
			if (upm) carry_start = __shfl_sync(0xffffffffu, pos, 31 - __clz(upm));
			depth = __shfl_sync(0xffffffffu, dp, 31);
		}
		if (lane == 0) {
			DSub s;
			uint32_t len = (uint32_t)(best >> 32);
			if (len > 0) {
				s.s_del = ((best_end - len) - clip) & 0x7fffffffu;
				s.e = best_end + clip;
				++remained;
			} else s.s_del = MAB_DEL_BIT, s.e = 0;
			sub[r] = s;
		}
	}
	if (lane == 0 && remained) atomicAdd(n_remained, (unsigned long long)remained);
}

// ---- shared-memory variant: one warp (<= SUBW_HITS hits) or one CTA (<= SUBC_HITS hits) per read ------------
// The endpoints of one read never leave the SM: load the group's hits (two 128-bit loads each), emit the keys into
// shared memory, bitonic-sort them there, sweep the depth with the same warp scan as k_sub_sweep.  Only reads
// with more hits than a CTA can hold go through the device-wide sort above.
constexpr int SUBW_WARPS = 8;
constexpr int SUBW_HITS = 256;               // per warp: 512 keys = 2 KB
constexpr int SUBC_HITS = 16384;             // per CTA: 32768 keys = 128 KB of dynamic shared memory (a power of two: the bitonic network pads up to it)

// depth sweep over n sorted keys in shared memory by one warp; returns the interval through *out (lane 0 writes)
__device__ __forceinline__ bool sub_sweep_smem(const uint32_t *key, uint32_t n, int min_dp, uint32_t clip, DSub *out, int lane)
{
	int depth = 0;
	uint32_t carry_start = 0, best_end = 0;
	unsigned long long best = 0;
	for (uint32_t c = 0; c < n; c += 32) {
		const uint32_t i = c + lane;
		const bool valid = i < n;
		const uint32_t k = valid ? key[i] : 0, pos = k >> 1;
		int delta = valid ? ((k & 1) ? -1 : 1) : 0, dp = delta;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, dp, o); if (lane >= o) dp += t; }
		dp += depth;
		const int old = dp - delta;
		const bool up = valid && old < min_dp && dp >= min_dp;
		const bool down = valid && old >= min_dp && dp < min_dp;
		const unsigned upm = __ballot_sync(0xffffffffu, up);
		const unsigned below = upm & ((1u << lane) - 1);
		const int src = below ? 31 - __clz(below) : 0;
		uint32_t st = __shfl_sync(0xffffffffu, pos, src);
		if (!below) st = carry_start;
		if (__any_sync(0xffffffffu, down)) { // an interval closes in this chunk (a handful per read): longest, earliest on ties
			unsigned long long cand = 0;
			if (down) cand = (unsigned long long)(pos - st) << 32 | (0xffffffffu - i);
			unsigned long long m = cand;
			#pragma unroll
			for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
			if (m > best && (m >> 32) != 0) {
				const unsigned who = __ballot_sync(0xffffffffu, down && cand == m);
				best = m;
				best_end = __shfl_sync(0xffffffffu, pos, __ffs(who) - 1);
			}
		}
		if (upm) carry_start = __shfl_sync(0xffffffffu, pos, 31 - __clz(upm));
		depth = __shfl_sync(0xffffffffu, dp, 31);
	}
	const uint32_t len = (uint32_t)(best >> 32);
	if (lane == 0) {
		DSub s;
		if (len > 0) s.s_del = ((best_end - len) - clip) & 0x7fffffffu, s.e = best_end + clip;
		else s.s_del = MAB_DEL_BIT, s.e = 0;
		*out = s;
	}
	return len > 0;
}

__device__ __forceinline__ bool sub_hit_keys(const DHit &h, uint32_t qid, float min_iden, uint32_t clip, uint32_t *ks, uint32_t *ke)
{
	const int ml = (int)(h.ml_rev & 0x7fffffffu), bl = (int)(h.bl_del & 0x7fffffffu);
	if (h.tn == qid || (float)ml < __fmul_rn((float)bl, min_iden)) return false;
	const uint32_t qs = (uint32_t)h.qns + clip, qe = h.qe - clip;
	if (!(qe > qs)) return false;
	*ks = qs << 1, *ke = qe << 1 | 1;
	return true;
}

// Bitonic network over 32*M keys held in registers, element e = 32*m + lane: exchanges at distance j < 32 are lane
// shuffles, distances >= 32 pair two registers of the same lane -- no shared-memory traffic and none of the 2-way bank
// conflicts the strided pair indexing has (ncu: 0.41 G conflicts per launch in the shared-memory version).
template <int M>
__device__ __forceinline__ void warp_bitonic_regs(uint32_t (&v)[M], const int lane)
{
	#pragma unroll
	for (int k = 2; k <= 32 * M; k <<= 1) {
		#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			if (j >= 32) {
				#pragma unroll
				for (int m = 0; m < M; ++m) {
					const int pm = m ^ (j >> 5);
					if (pm > m) {
						const bool asc = ((32 * m) & k) == 0; // k >= 64 here: the lane bits do not reach it
						const uint32_t x = v[m], y = v[pm];
						const uint32_t lo = min(x, y), hi = max(x, y);
						v[m] = asc ? lo : hi, v[pm] = asc ? hi : lo;
					}
				}
			} else {
				const bool low = (lane & j) == 0;
				#pragma unroll
				for (int m = 0; m < M; ++m) {
					const uint32_t y = __shfl_xor_sync(0xffffffffu, v[m], j);
					const bool asc = ((32 * m + lane) & k) == 0;
					v[m] = asc == low ? min(v[m], y) : max(v[m], y);
				}
			}
		}
	}
}

template <int M>
__device__ __forceinline__ void sub_sort_regs(uint32_t *key, const uint32_t n, const int lane)
{
	uint32_t v[M];
	#pragma unroll
	for (int m = 0; m < M; ++m) v[m] = 32u * m + lane < n ? key[32 * m + lane] : 0xffffffffu;
	warp_bitonic_regs<M>(v, lane);
	#pragma unroll
	for (int m = 0; m < M; ++m) key[32 * m + lane] = v[m];
}

__global__ void __launch_bounds__(SUBW_WARPS * 32)
k_sub_warp(const DHit *__restrict__ a, const uint64_t *__restrict__ grp, uint32_t n_seq, int min_dp, float min_iden, uint32_t clip,
           DSub *sub, uint32_t *big_list, unsigned long long *scal, int smem_sort)
{
	__shared__ uint32_t s_key[SUBW_WARPS][2 * SUBW_HITS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t *key = s_key[warp];
	unsigned remained = 0;
	for (uint32_t r = blockIdx.x * SUBW_WARPS + warp; r < n_seq; r += gridDim.x * SUBW_WARPS) {
		const uint64_t g = grp[r];
		const uint32_t end = (uint32_t)g, first = (uint32_t)(g >> 32);
		if (end == 0) continue;
		const uint32_t cnt = end - first;
		if (cnt > SUBW_HITS) { if (lane == 0) big_list[atomicAdd(scal + SC_BIG, 1ull)] = r; continue; }
		uint32_t n = 0; // keys emitted so far (uniform)
		for (uint32_t c = 0; c < cnt; c += 32) {
			uint32_t ks = 0, ke = 0;
			bool ok = false;
			if (c + lane < cnt) ok = sub_hit_keys(ld_hit(a + first + c + lane), r, min_iden, clip, &ks, &ke);
			const unsigned m = __ballot_sync(0xffffffffu, ok);
			if (ok) { const uint32_t p = n + 2 * __popc(m & ((1u << lane) - 1)); key[p] = ks, key[p + 1] = ke; }
			n += 2 * __popc(m);
		}
		uint32_t np = 32; while (np < n) np <<= 1;
		__syncwarp();
		if (!smem_sort) { // keys of the read sorted in registers (np is uniform across the warp)
			switch (np) {
				case 32: sub_sort_regs<1>(key, n, lane); break;
				case 64: sub_sort_regs<2>(key, n, lane); break;
				case 128: sub_sort_regs<4>(key, n, lane); break;
				case 256: sub_sort_regs<8>(key, n, lane); break;
				default: sub_sort_regs<16>(key, n, lane); break;
			}
			__syncwarp();
			remained += sub_sweep_smem(key, n, min_dp, clip, sub + r, lane);
			__syncwarp();
			continue;
		}
		for (uint32_t i = n + lane; i < np; i += 32) key[i] = 0xffffffffu;
		__syncwarp();
		// bitonic network; j = 1 << lj is a power of two, so the pair index comes from shifts (a 32-bit division per
		// compare-exchange made this loop ~90 % of the kernel's instructions: ncu, profiles/r01_launches_c3_summary.txt)
		for (uint32_t k = 2, lk = 1; k <= np; k <<= 1, ++lk)
			for (uint32_t lj = lk; lj-- > 0;) {
				const uint32_t j = 1u << lj;
				for (uint32_t t = lane; t < np / 2; t += 32) {
					const uint32_t lo = ((t >> lj) << (lj + 1)) | (t & (j - 1)), hi = lo + j;
					const uint32_t x = key[lo], y = key[hi];
					const bool asc = (lo & k) == 0;
					if ((x > y) == asc) key[lo] = y, key[hi] = x;
				}
				__syncwarp();
			}
		remained += sub_sweep_smem(key, n, min_dp, clip, sub + r, lane);
		__syncwarp();
	}
	if (lane == 0 && remained) atomicAdd(scal + SC_COUNT, (unsigned long long)remained);
}

// Large groups (hot spots: thousands of hits on one read).  The depth sweep only depends on HOW MANY intervals start and end at
// every coordinate, so a read whose coordinates fit the counter array is not sorted at all: one pass counts starts / ends per
// coordinate (16 + 16 bits in one shared-memory word; a group has <= 16 384 hits), a block scan turns the counts into depths, and the
// up / down crossings of min_dp are read off per coordinate -- starts before ends at a coordinate, exactly the order of the sorted keys
// (hit.c:137-149).  O(hits + coordinates) instead of the 120-stage bitonic network over 32 768 keys, which is kept for reads whose
// coordinates reach 32 768 or beyond.
constexpr uint32_t SUBC_COORDS = 2 * SUBC_HITS;      // counters that fit the CTA's 128 KB

__device__ __forceinline__ bool sub_count_sweep(uint32_t *cnt, uint32_t n_coord, int min_dp, uint32_t clip, DSub *out)
{	// cnt[p] = starts(p) | ends(p) << 16; all 512 threads of the CTA; returns (to every thread) whether an interval was kept
	__shared__ int s_part[512];
	__shared__ uint32_t s_up[512];
	__shared__ unsigned long long s_best[512];
	const int tid = threadIdx.x;
	const uint32_t per = (n_coord + 511) / 512, p0 = tid * per, p1 = p0 + per < n_coord ? p0 + per : n_coord;
	int delta = 0;
	for (uint32_t p = p0; p < p1; ++p) { const uint32_t c = cnt[p]; delta += (int)(c & 0xffffu) - (int)(c >> 16); }
	s_part[tid] = delta;
	__syncthreads();
	for (int o = 1; o < 512; o <<= 1) { // inclusive scan of the per-thread depth changes
		const int t = tid >= o ? s_part[tid - o] : 0;
		__syncthreads();
		s_part[tid] += t;
		__syncthreads();
	}
	const int depth0 = s_part[tid] - delta;            // depth before coordinate p0
	uint32_t last_up = 0xffffffffu;                    // pass 1: the last coordinate of my range where the depth rises through min_dp
	{
		int dp = depth0;
		for (uint32_t p = p0; p < p1; ++p) {
			const uint32_t c = cnt[p];
			const int mid = dp + (int)(c & 0xffffu);
			if (dp < min_dp && mid >= min_dp) last_up = p;
			dp = mid - (int)(c >> 16);
		}
	}
	s_up[tid] = last_up;
	__syncthreads();
	for (int o = 1; o < 512; o <<= 1) { // "latest crossing at or before my range": positions ascend with the thread index, NONE = 0xffffffff
		const uint32_t t = tid >= o ? s_up[tid - o] : 0xffffffffu;
		__syncthreads();
		if (s_up[tid] == 0xffffffffu) s_up[tid] = t;
		__syncthreads();
	}
	uint32_t start = tid ? s_up[tid - 1] : 0xffffffffu; // the interval that is open when my range begins started here
	if (start == 0xffffffffu) start = 0;               // (the reference's `start` is 0 until the first crossing)
	unsigned long long best = 0;                       // pass 2: intervals closing in my range: longest, earliest on ties
	{
		int dp = depth0;
		for (uint32_t p = p0; p < p1; ++p) {
			const uint32_t c = cnt[p];
			const int mid = dp + (int)(c & 0xffffu);
			if (dp < min_dp && mid >= min_dp) start = p;
			dp = mid - (int)(c >> 16);
			if (mid >= min_dp && dp < min_dp) {
				const unsigned long long cand = (unsigned long long)(p - start) << 32 | (0xffffffffu - p);
				if ((cand >> 32) != 0 && cand > best) best = cand;
			}
		}
	}
	s_best[tid] = best;
	__syncthreads();
	for (int o = 256; o; o >>= 1) {
		if (tid < o && s_best[tid + o] > s_best[tid]) s_best[tid] = s_best[tid + o];
		__syncthreads();
	}
	best = s_best[0];
	const uint32_t len = (uint32_t)(best >> 32), end = 0xffffffffu - (uint32_t)best;
	if (tid == 0) {
		DSub sres;
		if (len > 0) sres.s_del = ((end - len) - clip) & 0x7fffffffu, sres.e = end + clip;
		else sres.s_del = MAB_DEL_BIT, sres.e = 0;
		*out = sres;
	}
	__syncthreads();
	return len > 0;
}

__global__ void __launch_bounds__(512)
k_sub_cta(const DHit *__restrict__ a, const uint64_t *__restrict__ grp, const uint32_t *__restrict__ big_list, uint32_t n_big,
          int min_dp, float min_iden, uint32_t clip, DSub *sub, uint32_t *huge_list, unsigned long long *scal)
{
	extern __shared__ uint32_t c_key[];
	__shared__ uint32_t s_n, s_max;
	const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
	for (uint32_t b = blockIdx.x; b < n_big; b += gridDim.x) {
		const uint32_t r = big_list[b];
		const uint64_t g = grp[r];
		const uint32_t first = (uint32_t)(g >> 32), cnt = (uint32_t)g - first;
		if (cnt > SUBC_HITS) { if (tid == 0) huge_list[atomicAdd(scal + SC_AUX2, 1ull)] = r; continue; }
		if (tid == 0) s_n = 0, s_max = 0;
		__syncthreads();
		{ // largest coordinate any kept hit of the group touches
			uint32_t mx = 0;
			for (uint32_t c = tid; c < cnt; c += nt) {
				uint32_t ks, ke;
				if (sub_hit_keys(ld_hit(a + first + c), r, min_iden, clip, &ks, &ke)) mx = (ke >> 1) > mx ? (ke >> 1) : mx;
			}
			mx = __reduce_max_sync(0xffffffffu, mx);
			if (lane == 0 && mx) atomicMax(&s_max, mx);
		}
		__syncthreads();
		if (s_max < SUBC_COORDS) { // counting sweep
			const uint32_t n_coord = s_max + 1;
			for (uint32_t i = tid; i < n_coord; i += nt) c_key[i] = 0;
			__syncthreads();
			for (uint32_t c = tid; c < cnt; c += nt) {
				uint32_t ks, ke;
				if (sub_hit_keys(ld_hit(a + first + c), r, min_iden, clip, &ks, &ke)) { atomicAdd(&c_key[ks >> 1], 1u); atomicAdd(&c_key[ke >> 1], 1u << 16); }
			}
			__syncthreads();
			const bool kept = sub_count_sweep(c_key, n_coord, min_dp, clip, sub + r);
			if (tid == 0 && kept) atomicAdd(scal + SC_COUNT, 1ull);
			__syncthreads();
			continue;
		}
		for (uint32_t c = tid; c < cnt; c += nt) {
			uint32_t ks, ke;
			if (sub_hit_keys(ld_hit(a + first + c), r, min_iden, clip, &ks, &ke)) { const uint32_t p = atomicAdd(&s_n, 2u); c_key[p] = ks, c_key[p + 1] = ke; }
		}
		__syncthreads();
		const uint32_t n = s_n;
		uint32_t np = 32; while (np < n) np <<= 1;
		for (uint32_t i = n + tid; i < np; i += nt) c_key[i] = 0xffffffffu;
		__syncthreads();
		for (uint32_t k = 2, lk = 1; k <= np; k <<= 1, ++lk)
			for (uint32_t lj = lk; lj-- > 0;) {
				const uint32_t j = 1u << lj;
				for (uint32_t t = tid; t < np / 2; t += nt) {
					const uint32_t lo = ((t >> lj) << (lj + 1)) | (t & (j - 1)), hi = lo + j;
					const uint32_t x = c_key[lo], y = c_key[hi];
					const bool asc = (lo & k) == 0;
					if ((x > y) == asc) c_key[lo] = y, c_key[hi] = x;
				}
				__syncthreads();
			}
		if (tid < 32) {
			const bool kept = sub_sweep_smem(c_key, n, min_dp, clip, sub + r, lane);
			if (lane == 0 && kept) atomicAdd(scal + SC_COUNT, 1ull);
		}
		__syncthreads();
	}
}

__global__ void k_sub_copy_listed(const uint32_t *list, uint32_t n, const DSub *from, DSub *to, unsigned long long *n_remained)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		to[list[i]] = from[list[i]];
		if (!(from[list[i]].s_del & MAB_DEL_BIT)) atomicAdd(n_remained, 1ull);
	}
}

static uint64_t dh_sub_global(MabDev &d, const DHits &h, int min_dp, float min_iden, int end_clip, DSub *sub_out, const uint64_t *grp);

uint64_t dh_sub(MabDev &d, const DHits &h, int min_dp, float min_iden, int end_clip, DSub *sub_out)
{
	const uint32_t n_seq = h.n_seq;
	if (n_seq) MAB_CUDA(cudaMemsetAsync(sub_out, 0, (size_t)n_seq * sizeof(DSub), d.stream));
	if (h.n == 0 || n_seq == 0) return 0;
	if (h.n >= (1ull << 31)) { fprintf(stderr, "[E::miniasm_b200] more than 2^31 hits on one GPU\n"); exit(73); }
	uint64_t *grp = mab_alloc<uint64_t>(d, n_seq);
	uint32_t *big = mab_alloc<uint32_t>(d, n_seq);
	MAB_CUDA(cudaMemsetAsync(grp, 0, (size_t)n_seq * 8, d.stream));
	MAB_CUDA(cudaMemsetAsync(d.d_scal, 0, 8 * sizeof(unsigned long long), d.stream));
	MAB_LAUNCH(d, k_group_bounds, mab_grid(h.n, 256), 256, 0, h.a, h.n, (uint32_t*)grp);
	unsigned grid = (n_seq + SUBW_WARPS - 1) / SUBW_WARPS;
	if (grid > 148u * 32u) grid = 148u * 32u;
	static const int smem_sort = getenv("MAB_SUB_SMEM_SORT") && atoi(getenv("MAB_SUB_SMEM_SORT")) != 0; // 1: the shared-memory network
	MAB_LAUNCH(d, k_sub_warp, grid, SUBW_WARPS * 32, 0, h.a, grp, n_seq, min_dp, min_iden, (uint32_t)end_clip, sub_out, big, d.d_scal, smem_sort);
	uint32_t n_big = (uint32_t)d.get_scal(SC_BIG);
	if (n_big) {
		const size_t smem = (size_t)SUBC_HITS * 2 * 4;
		MAB_CUDA(cudaFuncSetAttribute(k_sub_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); // per device: set on every use
		uint32_t *huge = mab_alloc<uint32_t>(d, n_big);
		MAB_LAUNCH(d, k_sub_cta, n_big < 148u * 2 ? n_big : 148u * 2, 512, smem, h.a, grp, big, n_big, min_dp, min_iden, (uint32_t)end_clip, sub_out, huge, d.d_scal);
		uint32_t n_huge = (uint32_t)d.get_scal(SC_AUX2);
		if (n_huge) { // reads with more hits than a CTA can sort: device-wide sort of everything, keep only their rows
			DSub *tmp = mab_alloc<DSub>(d, n_seq);
			unsigned long long saved = d.get_scal(SC_COUNT);
			dh_sub_global(d, h, min_dp, min_iden, end_clip, tmp, grp);
			MAB_CUDA(cudaMemcpyAsync(d.d_scal + SC_COUNT, &saved, 8, cudaMemcpyHostToDevice, d.stream));
			MAB_LAUNCH(d, k_sub_copy_listed, mab_grid(n_huge, 128), 128, 0, huge, n_huge, tmp, sub_out, d.d_scal + SC_COUNT);
			d.sync();
			d.free(tmp);
		}
		d.free(huge);
	}
	uint64_t n_remained = d.get_scal(SC_COUNT);
	d.free(grp); d.free(big);
	{ unsigned long long v[1] = { n_remained }; sum_ranks(v, 1); n_remained = v[0]; }
	if (ma_verbose_dev >= 3)
		fprintf(stderr, "[M::%s::%s] %ld query sequences remain after sub\n", "ma_hit_sub", sys_timestamp(), (long)n_remained);
	return n_remained;
}

static uint64_t dh_sub_global(MabDev &d, const DHits &h, int min_dp, float min_iden, int end_clip, DSub *sub_out, const uint64_t *grp)
{
	const uint32_t n_seq = h.n_seq;
	if (n_seq) MAB_CUDA(cudaMemsetAsync(sub_out, 0, (size_t)n_seq * sizeof(DSub), d.stream));
	if (h.n == 0 || n_seq == 0) return 0;
	if (h.n >= (1ull << 31)) { fprintf(stderr, "[E::miniasm_b200] more than 2^31 hits on one GPU\n"); exit(73); }
	size_t nk = 2 * h.n;
	uint64_t *ka = mab_alloc<uint64_t>(d, nk), *kb = mab_alloc<uint64_t>(d, nk);
	MAB_LAUNCH(d, k_sub_keys, mab_grid(h.n, 256), 256, 0, h.a, h.n, min_iden, (uint32_t)end_clip, ka);
	cub::DoubleBuffer<uint64_t> dk(ka, kb);
	size_t tb = 0;
	int end_bit = 33 + (int)bits_for(n_seq - 1);
	cub::DeviceRadixSort::SortKeys(nullptr, tb, dk, (int64_t)nk, 0, end_bit, d.stream);
	void *tmp = d.tmp(tb);
	cub::DeviceRadixSort::SortKeys(tmp, tb, dk, (int64_t)nk, 0, end_bit, d.stream);
	++d.n_lib;
	d.zero_scal(SC_COUNT);
	MAB_LAUNCH(d, k_sub_sweep, mab_grid((size_t)n_seq * 32, 256), 256, 0, dk.Current(), grp, n_seq, min_dp, (uint32_t)end_clip, sub_out, d.d_scal + SC_COUNT);
	uint64_t n_remained = d.get_scal(SC_COUNT);
	d.free(ka); d.free(kb);
	return n_remained;
}

// ma_hit_cut (hit.c:162-193): the clipping rule itself is mab_cut_hit in hit2arc.cuh (shared with the CPU-tier check of the
// conversion rules, tests/hostsim/hit_host.cpp)
__global__ void k_cut(DHit *a, size_t n, const DSub *__restrict__ reg, int min_span, uint8_t *flag)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit p = ld_hit_rw(a + i);
		const bool keep = mab_cut_hit(p, reg[p.qns >> 32], reg[p.tn], min_span);
		if (keep) st_hit(a + i, p);
		flag[i] = keep;
	}
}

// ma_hit_cut followed by ma_hit_flt on the same interval table (main.c:123-125) in one sweep and one compaction
__global__ void k_cut_flt(DHit *a, size_t n, const DSub *__restrict__ sub, int min_span, int max_hang, int min_ovlp, uint8_t *flag,
                          unsigned long long *n_cut, unsigned long long *tot_dp)
{
	unsigned long long dp = 0;
	unsigned cut = 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit p = ld_hit_rw(a + i);
		const DSub sq = sub[p.qns >> 32], st = sub[p.tn];
		bool keep = mab_cut_hit(p, sq, st, min_span);
		if (keep) {
			++cut;
			DArc t;
			const uint32_t ql = sq.e - sq.s_del, tl = st.e - st.s_del;
			const int r = mab_hit2arc(p, (int)ql, (int)tl, max_hang, .5f, min_ovlp, &t);
			keep = r >= 0 || r == MAB_HT_QCONT || r == MAB_HT_TCONT;
			if (keep) { dp += r >= 0 ? (uint32_t)r : r == MAB_HT_QCONT ? ql : tl; st_hit(a + i, p); }
		}
		flag[i] = keep;
	}
	typedef cub::BlockReduce<unsigned long long, 256> BR;
	__shared__ typename BR::TempStorage ts;
	unsigned long long s = BR(ts).Sum(dp);
	__syncthreads();
	unsigned long long c = BR(ts).Sum((unsigned long long)cut);
	if (threadIdx.x == 0) { if (s) atomicAdd(tot_dp, s); if (c) atomicAdd(n_cut, c); }
}

// second-round ma_hit_cut fused with pass 1 of ma_hit_contained (hit.c:231-236): clip, then classify the clipped hit
// against the merged interval table (its lengths equal those of the round-2 table) and raise containment flags
__global__ void k_cut_cont_mark(DHit *a, size_t n, const DSub *__restrict__ sub2, DSub *sub, int min_span, HitArcParams hp,
                                uint8_t *used, uint8_t *flag, unsigned long long *n_cut)
{
	unsigned cut = 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit p = ld_hit_rw(a + i);
		const uint32_t q = (uint32_t)(p.qns >> 32), t = p.tn;
		const DSub rq = sub2[q], rt = sub2[t];
		const bool keep = mab_cut_hit(p, rq, rt, min_span);
		if (keep) {
			++cut;
			st_hit(a + i, p);
			DArc tmp;
			const int r = mab_hit2arc(p, (int)(rq.e - rq.s_del), (int)(rt.e - rt.s_del), hp.max_hang, hp.int_frac, hp.min_ovlp, &tmp);
			if (r == MAB_HT_QCONT) atomicOr(&sub[q].s_del, MAB_DEL_BIT);
			else if (r == MAB_HT_TCONT) atomicOr(&sub[t].s_del, MAB_DEL_BIT);
			used[q] = 1, used[t] = 1;
		}
		flag[i] = keep;
	}
	cut = __reduce_add_sync(0xffffffffu, cut);
	if ((threadIdx.x & 31) == 0 && cut) atomicAdd(n_cut, (unsigned long long)cut);
}

__global__ void k_cont_apply2(DHit *a, size_t n, const int32_t *__restrict__ map, uint8_t *flag)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		bool keep = flag[i];
		if (keep) {
			uint64_t qns = a[i].qns;
			int32_t qn = map[qns >> 32], tn = map[a[i].tn];
			keep = qn >= 0 && tn >= 0;
			if (keep) { a[i].qns = (uint64_t)(uint32_t)qn << 32 | (uint32_t)qns; a[i].tn = (uint32_t)tn; }
		}
		flag[i] = keep;
	}
}

__global__ void k_flt_len(const DHit *a, size_t n, const DSub *__restrict__ sub, unsigned long long *tot_len);

size_t dh_cut_flt(MabDev &d, DHits &h, const DSub *sub, int min_span, int max_hang, int min_ovlp, float *cov)
{
	unsigned long long tot_dp = 0, tot_len = 0, n_cut = 0;
	if (h.n) {
		uint8_t *flag = mab_alloc<uint8_t>(d, h.n);
		d.zero_scal(SC_AUX, 2);
		d.zero_scal(SC_COUNT, 1);
		MAB_LAUNCH(d, k_cut_flt, mab_grid(h.n, 256), 256, 0, h.a, h.n, sub, min_span, max_hang, min_ovlp, flag, d.d_scal + SC_COUNT, d.d_scal + SC_AUX);
		select_hits(d, h, flag);
		d.free(flag);
		if (h.n) MAB_LAUNCH(d, k_flt_len, mab_grid(h.n, 256), 256, 0, h.a, h.n, sub, d.d_scal + SC_AUX2);
		tot_dp = d.get_scal(SC_AUX), tot_len = d.h_scal[SC_AUX2], n_cut = d.h_scal[SC_COUNT];
	}
	unsigned long long gv[4] = { tot_dp, tot_len, n_cut, (unsigned long long)h.n };
	sum_ranks(gv, 4);
	*cov = (float)((double)gv[0] / gv[1]);
	if (ma_verbose_dev >= 3) {
		fprintf(stderr, "[M::%s::%s] %ld hits remain after cut\n", "ma_hit_cut", sys_timestamp(), (long)gv[2]);
		fprintf(stderr, "[M::%s::%s] %ld hits remain after filtering; crude coverage after filtering: %.2f\n", "ma_hit_flt", sys_timestamp(), (long)gv[3], *cov);
	}
	return h.n;
}

size_t dh_cut(MabDev &d, DHits &h, const DSub *reg, int min_span)
{
	if (h.n) {
		uint8_t *flag = mab_alloc<uint8_t>(d, h.n);
		MAB_LAUNCH(d, k_cut, mab_grid(h.n, 256), 256, 0, h.a, h.n, reg, min_span, flag);
		select_hits(d, h, flag);
		d.free(flag);
	}
	{ unsigned long long v[1] = { (unsigned long long)h.n }; sum_ranks(v, 1);
	  if (ma_verbose_dev >= 3) fprintf(stderr, "[M::%s::%s] %ld hits remain after cut\n", "ma_hit_cut", sys_timestamp(), (long)v[0]); }
	return h.n;
}

// ---------------------------------------------------------------------------------------------
// ma_hit_flt (hit.c:195-216)
// ---------------------------------------------------------------------------------------------
__global__ void k_flt(const DHit *a, size_t n, const DSub *__restrict__ sub, int max_hang, int min_ovlp, uint8_t *flag, unsigned long long *tot_dp)
{
	unsigned long long dp = 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit h = ld_hit(a + i);
		const DSub sq = sub[h.qns >> 32], st = sub[h.tn];
		bool keep = false;
		if (!((sq.s_del | st.s_del) & MAB_DEL_BIT)) {
			DArc t;
			const uint32_t ql = sq.e - sq.s_del, tl = st.e - st.s_del;
			int r = mab_hit2arc(h, (int)ql, (int)tl, max_hang, .5f, min_ovlp, &t);
			if (r >= 0 || r == MAB_HT_QCONT || r == MAB_HT_TCONT) {
				keep = true;
				dp += r >= 0 ? (uint32_t)r : r == MAB_HT_QCONT ? ql : tl;
			}
		}
		flag[i] = keep;
	}
	typedef cub::BlockReduce<unsigned long long, 256> BR;
	__shared__ typename BR::TempStorage ts;
	unsigned long long s = BR(ts).Sum(dp);
	if (threadIdx.x == 0 && s) atomicAdd(tot_dp, s);
}

__global__ void k_flt_len(const DHit *a, size_t n, const DSub *__restrict__ sub, unsigned long long *tot_len)
{
	unsigned long long len = 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint32_t q = (uint32_t)(a[i].qns >> 32);
		if (i == n - 1 || (uint32_t)(a[i + 1].qns >> 32) != q) len += sub[q].e - (sub[q].s_del & 0x7fffffffu);
	}
	typedef cub::BlockReduce<unsigned long long, 256> BR;
	__shared__ typename BR::TempStorage ts;
	unsigned long long s = BR(ts).Sum(len);
	if (threadIdx.x == 0 && s) atomicAdd(tot_len, s);
}

size_t dh_flt(MabDev &d, DHits &h, const DSub *sub, int max_hang, int min_ovlp, float *cov)
{
	unsigned long long tot_dp = 0, tot_len = 0;
	if (h.n) {
		uint8_t *flag = mab_alloc<uint8_t>(d, h.n);
		d.zero_scal(SC_AUX, 2);
		MAB_LAUNCH(d, k_flt, mab_grid(h.n, 256), 256, 0, h.a, h.n, sub, max_hang, min_ovlp, flag, d.d_scal + SC_AUX);
		select_hits(d, h, flag);
		d.free(flag);
		if (h.n) MAB_LAUNCH(d, k_flt_len, mab_grid(h.n, 256), 256, 0, h.a, h.n, sub, d.d_scal + SC_AUX2);
		tot_dp = d.get_scal(SC_AUX), tot_len = d.h_scal[SC_AUX2];
	}
	unsigned long long gv[3] = { tot_dp, tot_len, (unsigned long long)h.n };
	sum_ranks(gv, 3);
	*cov = (float)((double)gv[0] / gv[1]);
	if (ma_verbose_dev >= 3)
		fprintf(stderr, "[M::%s::%s] %ld hits remain after filtering; crude coverage after filtering: %.2f\n", "ma_hit_flt", sys_timestamp(), (long)gv[2], *cov);
	return h.n;
}

// ---------------------------------------------------------------------------------------------
// ma_sub_merge (hit.c:218-223): a.e = a.s + b.e; a.s += b.s  (31-bit s field; del of `a` is kept, del of `b` ignored)
// ---------------------------------------------------------------------------------------------
__global__ void k_sub_merge(uint32_t n, DSub *a, const DSub *b)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		DSub x = a[i], y = b[i];
		uint32_t s = x.s_del & 0x7fffffffu, del = x.s_del & MAB_DEL_BIT;
		x.e = s + y.e;
		x.s_del = ((s + (y.s_del & 0x7fffffffu)) & 0x7fffffffu) | del;
		a[i] = x;
	}
}

void dh_sub_merge(MabDev &d, uint32_t n_sub, DSub *a, const DSub *b)
{
	if (n_sub) MAB_LAUNCH(d, k_sub_merge, mab_grid(n_sub, 256), 256, 0, n_sub, a, b);
}

// ---------------------------------------------------------------------------------------------
// ma_hit_contained (hit.c:225-256)
// ---------------------------------------------------------------------------------------------
__global__ void k_cont_mark(const DHit *a, size_t n, DSub *sub, HitArcParams p, uint8_t *used)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit h = ld_hit(a + i);
		const uint32_t q = (uint32_t)(h.qns >> 32), t = h.tn;
		const DSub sq = sub[q], st = sub[t];
		DArc tmp;
		// lengths use the 31-bit s field only, so concurrent del-bit updates by other threads cannot change them
		int r = mab_hit2arc(h, (int)(sq.e - (sq.s_del & 0x7fffffffu)), (int)(st.e - (st.s_del & 0x7fffffffu)), p.max_hang, p.int_frac, p.min_ovlp, &tmp);
		if (r == MAB_HT_QCONT) atomicOr(&sub[q].s_del, MAB_DEL_BIT);
		else if (r == MAB_HT_TCONT) atomicOr(&sub[t].s_del, MAB_DEL_BIT);
		used[q] = 1, used[t] = 1;                   // ma_hit_mark_unused: a read is "used" if any of the n hits names it
	}
}

__global__ void k_cont_keep(uint32_t n_seq, const DSub *sub, const uint8_t *used, const uint8_t *seq_del, uint32_t *keep)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_seq; i += gridDim.x * blockDim.x)
		keep[i] = used[i] && !(sub[i].s_del & MAB_DEL_BIT) && !(seq_del && seq_del[i]);
}

__global__ void k_cont_map(uint32_t n_seq, const uint32_t *keep, const uint32_t *excl, int32_t *map, const DSub *sub, DSub *sub_out)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_seq; i += gridDim.x * blockDim.x) {
		if (keep[i]) { map[i] = (int32_t)excl[i]; sub_out[excl[i]] = sub[i]; }
		else map[i] = -1;
	}
}

__global__ void k_cont_apply(DHit *a, size_t n, const int32_t *__restrict__ map, uint8_t *flag)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint64_t qns = a[i].qns;
		int32_t qn = map[qns >> 32], tn = map[a[i].tn];
		bool keep = qn >= 0 && tn >= 0;
		if (keep) { a[i].qns = (uint64_t)(uint32_t)qn << 32 | (uint32_t)qns; a[i].tn = (uint32_t)tn; }
		flag[i] = keep;
	}
}

size_t dh_contained(MabDev &d, DHits &h, DSub *sub, const uint8_t *seq_del, const HitArcParams &p, int32_t *map_out, const DSub *cut_reg, int min_span,
                    const std::function<void(DSub*, uint8_t*, uint32_t)> *exchange)
{
	const uint32_t n_seq = h.n_seq;
	uint32_t n_new = 0;
	unsigned long long n_cut = 0;
	uint8_t *cut_flag = nullptr;
	if (n_seq) {
		uint8_t *used = mab_alloc<uint8_t>(d, n_seq);
		uint32_t *keep = mab_alloc<uint32_t>(d, n_seq), *excl = mab_alloc<uint32_t>(d, (size_t)n_seq + 1);
		DSub *sub2 = mab_alloc<DSub>(d, n_seq);
		MAB_CUDA(cudaMemsetAsync(used, 0, n_seq, d.stream));
		if (cut_reg && h.n) { // fused with the preceding ma_hit_cut: one sweep clips and classifies, one compaction at the end
			cut_flag = mab_alloc<uint8_t>(d, h.n);
			d.zero_scal(SC_COUNT, 1);
			MAB_LAUNCH(d, k_cut_cont_mark, mab_grid(h.n, 256), 256, 0, h.a, h.n, cut_reg, sub, min_span, p, used, cut_flag, d.d_scal + SC_COUNT);
		} else
		if (h.n) MAB_LAUNCH(d, k_cont_mark, mab_grid(h.n, 256), 256, 0, h.a, h.n, sub, p, used);
		if (exchange) (*exchange)(sub, used, n_seq); // sharded runs: OR the containment flags and the "used" marks of all ranks
		MAB_LAUNCH(d, k_cont_keep, mab_grid(n_seq, 256), 256, 0, n_seq, sub, used, seq_del, keep);
		size_t tb = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tb, keep, excl, (int)n_seq, d.stream);
		void *tmp = d.tmp(tb);
		cub::DeviceScan::ExclusiveSum(tmp, tb, keep, excl, (int)n_seq, d.stream);
		++d.n_lib;
		MAB_LAUNCH(d, k_cont_map, mab_grid(n_seq, 256), 256, 0, n_seq, keep, excl, map_out, sub, sub2);
		uint32_t last_keep, last_excl;
		MAB_CUDA(cudaMemcpyAsync(&last_keep, keep + n_seq - 1, 4, cudaMemcpyDeviceToHost, d.stream));
		MAB_CUDA(cudaMemcpyAsync(&last_excl, excl + n_seq - 1, 4, cudaMemcpyDeviceToHost, d.stream));
		d.sync();
		n_new = last_keep + last_excl;
		if (n_new) MAB_CUDA(cudaMemcpyAsync(sub, sub2, (size_t)n_new * sizeof(DSub), cudaMemcpyDeviceToDevice, d.stream));
		if (cut_flag) {
			n_cut = d.get_scal(SC_COUNT);
			MAB_LAUNCH(d, k_cont_apply2, mab_grid(h.n, 256), 256, 0, h.a, h.n, map_out, cut_flag);
			select_hits(d, h, cut_flag);
			d.free(cut_flag);
		} else if (h.n) {
			uint8_t *flag = mab_alloc<uint8_t>(d, h.n);
			MAB_LAUNCH(d, k_cont_apply, mab_grid(h.n, 256), 256, 0, h.a, h.n, map_out, flag);
			select_hits(d, h, flag);
			d.free(flag);
		}
		d.free(used); d.free(keep); d.free(excl); d.free(sub2);
	}
	h.n_seq = n_new;
	unsigned long long gv[2] = { (unsigned long long)n_cut, (unsigned long long)h.n };
	sum_ranks(gv, 2);
	if (cut_reg && ma_verbose_dev >= 3) fprintf(stderr, "[M::%s::%s] %ld hits remain after cut\n", "ma_hit_cut", sys_timestamp(), (long)gv[0]);
	if (ma_verbose_dev >= 3)
		fprintf(stderr, "[M::%s::%s] %d sequences and %ld hits remain after containment removal\n", "ma_hit_contained", sys_timestamp(), n_new, (long)gv[1]);
	return h.n;
}

// ---------------------------------------------------------------------------------------------
// ma_sg_gen (asm.c:9-39)
// ---------------------------------------------------------------------------------------------
__global__ void k_seq_set(uint32_t n_seq, const uint32_t *len, const uint8_t *del, uint32_t *seq, unsigned *max_len)
{
	unsigned mx = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_seq; i += gridDim.x * blockDim.x) {
		uint32_t l = len[i] & 0x7fffffffu;
		seq[i] = l | (del && del[i] ? MAB_DEL_BIT : 0);
		mx = l > mx ? l : mx;
	}
	mx = __reduce_max_sync(0xffffffffu, mx);
	if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_len, mx);
}

// one (key, value) column pair per hit: key = u << lb | len (or the sentinel 1 << (lb + vertex bits) when the hit yields no arc),
// value = ol:del << 32 | v.  Sorting the columns (stable) and dropping the sentinels equals asg.c's append + sort.
__global__ void k_sg_arcs(const DHit *a, size_t n, uint32_t *seq, HitArcParams p, uint32_t lb, uint64_t sentinel, uint64_t *key, uint64_t *val, unsigned long long *n_emit)
{
	unsigned cnt = 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit h = ld_hit(a + i);
		const uint32_t qn = (uint32_t)(h.qns >> 32);
		DArc t;
		t.ul = 0, t.v = 0, t.ol_del = 0;
		int r = mab_hit2arc(h, (int)(seq[qn] & 0x7fffffffu), (int)(seq[h.tn] & 0x7fffffffu), p.max_hang, p.int_frac, p.min_ovlp, &t);
		bool emit = false;
		if (r >= 0) {
			if (qn == h.tn) { // self match: only the palindromic artefact has an effect (asm.c:27-31)
				if ((uint32_t)h.qns == h.ts && h.qe == h.te && (h.ml_rev >> 31)) atomicOr(&seq[qn], MAB_DEL_BIT);
			} else emit = true;
		} else if (r == MAB_HT_QCONT) atomicOr(&seq[qn], MAB_DEL_BIT);
		key[i] = emit ? ((t.ul >> 32) << lb | (uint32_t)t.ul) : sentinel;
		val[i] = (uint64_t)t.ol_del << 32 | t.v;
		cnt += emit;
	}
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_emit, (unsigned long long)cnt);
}

// ---------------------------------------------------------------------------------------------
// ma_sg_gen without a device-wide sort (the default; MAB_SG_SEGSORT=0 or a case it declines -> the column sort above).
// Hits arrive grouped by query read, and an arc's source vertex is its hit's query, so asg.c's "append, then sort by
// ul" only has to order each read's arcs by (direction, length), hit order on ties -- a per-read problem of ~100
// elements.  Pass 1 classifies every hit (flag byte, group bounds, the deletion side effects of asm.c:27-33) and
// checks that query ids ascend; a scan of the flags gives each read its output offset; pass 2 re-classifies a read's
// hits (cheaper than storing 16 B per hit), sorts the keys ((dir << lb | len) << 9 | arc number) with the register
// network of ma_hit_sub and writes the read's arcs in place.  Reads with 257..8192 hits take a CTA and a shared-memory
// network; anything the scheme cannot take (unsorted ids, a read beyond 8192 hits, reads longer than 4 Mb) falls
// back to the column sort, so results never depend on the switch.
// ---------------------------------------------------------------------------------------------
constexpr int SGW_WARPS = 8, SGW_HITS = 256, SGW_IDX_BITS = 9;
constexpr int SGC_HITS = 8192;

__global__ void k_sg_classify(const DHit *a, size_t n, uint32_t *seq, HitArcParams p, uint8_t *emit_flag, uint32_t *g32, unsigned long long *scal)
{
	unsigned cnt = 0;
	bool bad = false;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		DHit h = ld_hit(a + i);
		const uint32_t qn = (uint32_t)(h.qns >> 32);
		DArc t;
		t.ul = 0, t.v = 0, t.ol_del = 0;
		const int r = mab_hit2arc(h, (int)(seq[qn] & 0x7fffffffu), (int)(seq[h.tn] & 0x7fffffffu), p.max_hang, p.int_frac, p.min_ovlp, &t);
		bool emit = false;
		if (r >= 0) {
			if (qn == h.tn) {
				if ((uint32_t)h.qns == h.ts && h.qe == h.te && (h.ml_rev >> 31)) atomicOr(&seq[qn], MAB_DEL_BIT);
			} else emit = true;
		} else if (r == MAB_HT_QCONT) atomicOr(&seq[qn], MAB_DEL_BIT);
		emit_flag[i] = emit;
		cnt += emit;
		const uint32_t prev = i ? (uint32_t)(a[i - 1].qns >> 32) : 0;
		if (i == 0 || prev != qn) g32[2 * (size_t)qn + 1] = (uint32_t)i;
		if (i == n - 1 || (uint32_t)(a[i + 1].qns >> 32) != qn) g32[2 * (size_t)qn] = (uint32_t)(i + 1);
		if (i && prev > qn) bad = true;
	}
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(scal + SC_COUNT, (unsigned long long)cnt);
	if (bad) scal[SC_AUX2] = 1ull;
}

struct U8ToU32 { __host__ __device__ __forceinline__ uint32_t operator()(uint8_t x) const { return x; } };

__global__ void __launch_bounds__(SGW_WARPS * 32)
k_sg_sort_warp(const DHit *__restrict__ a, const uint64_t *__restrict__ grp, const uint32_t *__restrict__ epos, const uint32_t *__restrict__ seq,
               uint32_t n_seq, HitArcParams p, uint32_t lb, DArc *__restrict__ out, uint32_t *big_list, unsigned long long *scal)
{
	__shared__ uint32_t s_key[SGW_WARPS][SGW_HITS], s_v[SGW_WARPS][SGW_HITS], s_ol[SGW_WARPS][SGW_HITS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t *key = s_key[warp], *pv = s_v[warp], *pol = s_ol[warp];
	const uint32_t len_mask = (1u << lb) - 1u; // lb <= 22 on this path
	for (uint32_t r = blockIdx.x * SGW_WARPS + warp; r < n_seq; r += gridDim.x * SGW_WARPS) {
		const uint64_t g = grp[r];
		const uint32_t end = (uint32_t)g, first = (uint32_t)(g >> 32);
		if (end == 0) continue;
		const uint32_t cnt = end - first;
		if (cnt > SGW_HITS) { if (lane == 0) big_list[atomicAdd(scal + SC_BIG, 1ull)] = r; continue; }
		const int ql = (int)(seq[r] & 0x7fffffffu);
		uint32_t n = 0; // arcs of this read so far (uniform)
		for (uint32_t c = 0; c < cnt; c += 32) {
			bool ok = false;
			DArc t;
			t.ul = 0, t.v = 0, t.ol_del = 0;
			if (c + lane < cnt) {
				const DHit h = ld_hit(a + first + c + lane);
				const int rr = mab_hit2arc(h, ql, (int)(seq[h.tn] & 0x7fffffffu), p.max_hang, p.int_frac, p.min_ovlp, &t);
				ok = rr >= 0 && h.tn != r;
			}
			const unsigned m = __ballot_sync(0xffffffffu, ok);
			if (ok) {
				const uint32_t k = n + __popc(m & ((1u << lane) - 1u));
				key[k] = ((((uint32_t)(t.ul >> 32) & 1u) << lb | (uint32_t)t.ul) << SGW_IDX_BITS) | k;
				pv[k] = t.v, pol[k] = t.ol_del;
			}
			n += __popc(m);
		}
		if (n == 0) continue;
		uint32_t np = 32; while (np < n) np <<= 1;
		__syncwarp();
		switch (np) {
			case 32: sub_sort_regs<1>(key, n, lane); break;
			case 64: sub_sort_regs<2>(key, n, lane); break;
			case 128: sub_sort_regs<4>(key, n, lane); break;
			default: sub_sort_regs<8>(key, n, lane); break;
		}
		__syncwarp();
		const uint32_t base = epos[first];
		for (uint32_t j = lane; j < n; j += 32) {
			const uint32_t k = key[j], src = k & ((1u << SGW_IDX_BITS) - 1u), kl = k >> SGW_IDX_BITS;
			*reinterpret_cast<uint4*>(out + base + j) = make_uint4(kl & len_mask, r << 1 | kl >> lb, pv[src], pol[src]);
		}
		__syncwarp();
	}
}

__global__ void __launch_bounds__(512)
k_sg_sort_cta(const DHit *__restrict__ a, const uint64_t *__restrict__ grp, const uint32_t *__restrict__ epos, const uint32_t *__restrict__ seq,
              const uint32_t *__restrict__ big_list, uint32_t n_big, HitArcParams p, DArc *__restrict__ out, unsigned long long *scal)
{
	extern __shared__ __align__(16) unsigned char sg_smem[];
	uint64_t *key = reinterpret_cast<uint64_t*>(sg_smem);         // (dir << 31 | len) << 32 | arc number
	uint32_t *pv = reinterpret_cast<uint32_t*>(key + SGC_HITS), *pol = pv + SGC_HITS;
	__shared__ uint32_t s_n;
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	for (uint32_t b = blockIdx.x; b < n_big; b += gridDim.x) {
		const uint32_t r = big_list[b];
		const uint64_t g = grp[r];
		const uint32_t first = (uint32_t)(g >> 32), cnt = (uint32_t)g - first;
		if (cnt > SGC_HITS) { if (tid == 0) atomicAdd(scal + SC_AUX, 1ull); continue; }
		if (tid == 0) s_n = 0;
		__syncthreads();
		const uint32_t base = epos[first];
		const int ql = (int)(seq[r] & 0x7fffffffu);
		for (uint32_t c = tid; c < cnt; c += nt) {
			const DHit h = ld_hit(a + first + c);
			DArc t;
			t.ul = 0, t.v = 0, t.ol_del = 0;
			const int rr = mab_hit2arc(h, ql, (int)(seq[h.tn] & 0x7fffffffu), p.max_hang, p.int_frac, p.min_ovlp, &t);
			if (rr >= 0 && h.tn != r) {
				const uint32_t k = epos[first + c] - base; // arcs of this read emitted before hit c
				key[k] = (uint64_t)((((uint32_t)(t.ul >> 32) & 1u) << 31) | (uint32_t)t.ul) << 32 | k;
				pv[k] = t.v, pol[k] = t.ol_del;
				atomicAdd(&s_n, 1u);
			}
		}
		__syncthreads();
		const uint32_t n = s_n;
		uint32_t np = 2; while (np < n) np <<= 1;
		for (uint32_t i = n + tid; i < np; i += nt) key[i] = ~0ull;
		__syncthreads();
		for (uint32_t k = 2, lk = 1; k <= np; k <<= 1, ++lk)
			for (uint32_t lj = lk; lj-- > 0;) {
				const uint32_t j = 1u << lj;
				for (uint32_t t = tid; t < np / 2; t += nt) {
					const uint32_t lo = ((t >> lj) << (lj + 1)) | (t & (j - 1)), hi = lo + j;
					const uint64_t x = key[lo], y = key[hi];
					const bool asc = (lo & k) == 0;
					if ((x > y) == asc) key[lo] = y, key[hi] = x;
				}
				__syncthreads();
			}
		for (uint32_t j = tid; j < n; j += nt) {
			const uint64_t k = key[j];
			const uint32_t src = (uint32_t)k, kl = (uint32_t)(k >> 32);
			*reinterpret_cast<uint4*>(out + base + j) = make_uint4(kl & 0x7fffffffu, r << 1 | kl >> 31, pv[src], pol[src]);
		}
		__syncthreads();
	}
}

// true: g.arc holds the sorted arcs.  false: nothing usable was produced, the caller runs the column sort.
static bool sg_emit_segmented(MabDev &d, const DHits &h, const HitArcParams &p, uint32_t lb, DGraph &g)
{
	const size_t n = h.n;
	const uint32_t n_seq = h.n_seq;
	if (lb + 1 + SGW_IDX_BITS > 32 || n >= (1ull << 31)) return false;
	uint8_t *ef = mab_alloc<uint8_t>(d, n);
	uint32_t *epos = mab_alloc<uint32_t>(d, n), *big = mab_alloc<uint32_t>(d, n_seq);
	uint64_t *grp = mab_alloc<uint64_t>(d, n_seq);
	MAB_CUDA(cudaMemsetAsync(grp, 0, (size_t)n_seq * 8, d.stream));
	d.zero_scal(SC_COUNT, 5); // SC_COUNT .. SC_AUX2
	MAB_LAUNCH(d, k_sg_classify, mab_grid(n, 256), 256, 0, h.a, n, g.seq, p, ef, (uint32_t*)grp, d.d_scal);
	const uint32_t n_arc = (uint32_t)d.get_scal(SC_COUNT);
	bool ok = d.h_scal[SC_AUX2] == 0; // query ids ascend
	if (ok) {
		cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> in(ef, U8ToU32());
		size_t tb = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tb, in, epos, (int64_t)n, d.stream);
		void *tmp = d.tmp(tb);
		cub::DeviceScan::ExclusiveSum(tmp, tb, in, epos, (int64_t)n, d.stream);
		++d.n_lib;
		dg_reserve(d, g, n_arc ? n_arc : 1);
		unsigned grid = (n_seq + SGW_WARPS - 1) / SGW_WARPS;
		if (grid > 148u * 32u) grid = 148u * 32u;
		MAB_LAUNCH(d, k_sg_sort_warp, grid, SGW_WARPS * 32, 0, h.a, grp, epos, g.seq, n_seq, p, lb, g.arc, big, d.d_scal);
		const uint32_t n_big = (uint32_t)d.get_scal(SC_BIG);
		if (n_big) {
				const size_t smem = (size_t)SGC_HITS * 16;
			MAB_CUDA(cudaFuncSetAttribute(k_sg_sort_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			MAB_LAUNCH(d, k_sg_sort_cta, n_big < 148u ? n_big : 148u, 512, smem, h.a, grp, epos, g.seq, big, n_big, p, g.arc, d.d_scal);
			if (d.get_scal(SC_AUX) != 0) ok = false; // a read with more hits than a CTA sorts
		}
	}
	d.free(ef); d.free(epos); d.free(big); d.free(grp);
	if (ok) g.len_bits = lb, g.n_arc = n_arc, g.is_srt = true, g.has_idx = false;
	return ok;
}

void dh_sg_gen(MabDev &d, const DHits &h, const uint32_t *len, const uint8_t *del, const HitArcParams &p, DGraph &g)
{
	dh_sg_emit(d, h, len, del, p, g);
	dg_cleanup(d, g);
	if (MAB_V(1)) fprintf(stderr, "[M::%s] read %d arcs\n", "ma_sg_gen", g.n_arc);
}

void dh_sg_emit(MabDev &d, const DHits &h, const uint32_t *len, const uint8_t *del, const HitArcParams &p, DGraph &g)
{
	const uint32_t n_seq = h.n_seq;
	dg_set_nseq(d, g, n_seq);
	g.n_arc = 0, g.is_srt = false, g.is_symm = false, g.has_idx = false;
	d.zero_scal(SC_AUX);
	d.zero_scal(SC_COUNT);
	unsigned *d_max = (unsigned*)(d.d_scal + SC_AUX);
	if (n_seq) MAB_LAUNCH(d, k_seq_set, mab_grid(n_seq, 256), 256, 0, n_seq, len, del, g.seq, d_max);
	const unsigned mx = (unsigned)(d.get_scal(SC_AUX) & 0xffffffffu); // an arc is never longer than its source read (miniasm.h:97-98)
	const uint32_t lb = bits_for(mx);
	if (h.n) {
		if (h.n >= (1ull << 31)) { fprintf(stderr, "[E::miniasm_b200] more than 2^31 arcs on one GPU\n"); exit(73); }
		static const bool seg_sort = !(getenv("MAB_SG_SEGSORT") && atoi(getenv("MAB_SG_SEGSORT")) == 0); // default on; 0 = device-wide column sort
		if (seg_sort && sg_emit_segmented(d, h, p, lb, g)) return;
		if (seg_sort) d.zero_scal(SC_COUNT); // the attempt counted the arcs already
		const uint64_t sentinel = 1ull << (lb + bits_for((uint64_t)n_seq * 2 - 1));
		uint64_t *ka = mab_alloc<uint64_t>(d, h.n), *kb = mab_alloc<uint64_t>(d, h.n), *va = mab_alloc<uint64_t>(d, h.n), *vb = mab_alloc<uint64_t>(d, h.n);
		// seq lengths are read while other threads may set del bits: lengths are masked, so this is benign
		MAB_LAUNCH(d, k_sg_arcs, mab_grid(h.n, 256), 256, 0, h.a, h.n, g.seq, p, lb, sentinel, ka, va, d.d_scal + SC_COUNT);
		const uint32_t n_arc = (uint32_t)d.get_scal(SC_COUNT);
		dg_build_sorted(d, g, ka, va, kb, vb, (uint32_t)h.n, n_arc, lb, true);
		d.free(ka); d.free(kb); d.free(va); d.free(vb);
	} else { dg_reserve(d, g, 1); g.is_srt = true; }
}
