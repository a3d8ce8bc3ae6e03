// mab_dev.cu -- per-GPU runtime object: stream, stream-ordered allocator, CUB scratch, scalar mailbox.
#include "mab_common.cuh"

void MabDev::init(int dev)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		fprintf(stderr, "[E::miniasm_b200] no usable CUDA device (%s); this library has no CPU path\n",
				e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		exit(71);
	}
	if (dev < 0 || dev >= n) {
		fprintf(stderr, "[E::miniasm_b200] device %d out of range (have %d)\n", dev, n);
		exit(71);
	}
	device = dev;
	MAB_CUDA(cudaSetDevice(dev));
	MAB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
	MAB_CUDA(cudaMalloc(&d_scal, 64 * sizeof(unsigned long long)));
	MAB_CUDA(cudaMemset(d_scal, 0, 64 * sizeof(unsigned long long)));
	MAB_CUDA(cudaMallocHost(&h_scal, 64 * sizeof(unsigned long long)));
}

void MabDev::destroy()
{
	if (!stream) return;
	MAB_CUDA(cudaSetDevice(device));
	MAB_CUDA(cudaStreamSynchronize(stream));
	MAB_CUDA(cudaStreamSynchronize(stream));
	arena.release_all();
	MAB_CUDA(cudaFree(d_scal));
	MAB_CUDA(cudaFreeHost(h_scal));
	if (copy_stream) MAB_CUDA(cudaStreamDestroy(copy_stream));
	copy_stream = nullptr;
	MAB_CUDA(cudaStreamDestroy(stream));
	stream = nullptr; cub_tmp = nullptr; cub_tmp_bytes = 0; d_scal = h_scal = nullptr;
}

void *MabArena::alloc(size_t bytes)
{
	bytes = (bytes + 511) & ~(size_t)511;
	if (bytes == 0) bytes = 512;
	// best fit among the free blocks (few dozen at most)
	auto best = free_blk.end();
	for (auto it = free_blk.begin(); it != free_blk.end(); ++it)
		if (it->second >= bytes && (best == free_blk.end() || it->second < best->second)) best = it;
	if (best == free_blk.end()) {
		size_t seg = bytes > ((size_t)256 << 20) ? bytes : ((size_t)256 << 20);
		char *base = nullptr;
		cudaError_t e = cudaMalloc(&base, seg);
		if (e != cudaSuccess) {
			fprintf(stderr, "[E::miniasm_b200] device allocation of %zu bytes failed (%zu already reserved): %s\n", seg, reserved, cudaGetErrorString(e));
			exit(72);
		}
		segs.push_back(Seg{base, seg});
		reserved += seg;
		best = free_blk.emplace(base, seg).first;
	}
	char *p = best->first;
	size_t sz = best->second;
	free_blk.erase(best);
	if (sz > bytes) free_blk.emplace(p + bytes, sz - bytes);
	live.emplace(p, bytes);
	in_use += bytes;
	if (in_use > peak) peak = in_use;
	return p;
}

void MabArena::free(void *ptr)
{
	if (!ptr) return;
	char *p = (char*)ptr;
	auto it = live.find(p);
	if (it == live.end()) { fprintf(stderr, "[E::miniasm_b200] arena: free of an unknown pointer\n"); exit(72); }
	size_t sz = it->second;
	live.erase(it);
	in_use -= sz;
	// coalesce with the neighbours when they belong to the same segment (segments are never adjacent by construction
	// of the check below: a block only merges if it ends exactly where the next one starts AND both lie in one segment)
	auto seg_of = [&](char *q) -> const Seg* { for (auto &s : segs) if (q >= s.base && q < s.base + s.size) return &s; return nullptr; };
	const Seg *sg = seg_of(p);
	auto nxt = free_blk.lower_bound(p);
	if (nxt != free_blk.end() && nxt->first == p + sz && seg_of(nxt->first) == sg) { sz += nxt->second; nxt = free_blk.erase(nxt); }
	if (nxt != free_blk.begin()) {
		auto prv = std::prev(nxt);
		if (prv->first + prv->second == p && seg_of(prv->first) == sg) { prv->second += sz; return; }
	}
	free_blk.emplace(p, sz);
}

void MabArena::release_all()
{
	for (auto &s : segs) cudaFree(s.base);
	segs.clear(); free_blk.clear(); live.clear();
	reserved = in_use = 0;
}

void *MabDev::alloc(size_t bytes) { return arena.alloc(bytes); }
void MabDev::free(void *p) { arena.free(p); }

void *MabDev::tmp(size_t bytes)
{
	if (bytes > cub_tmp_bytes) {
		if (cub_tmp) free(cub_tmp);
		cub_tmp_bytes = bytes + (bytes >> 2) + 256;
		cub_tmp = alloc(cub_tmp_bytes);
	}
	return cub_tmp;
}

void MabDev::zero_scal(int i, int n)
{
	MAB_CUDA(cudaMemsetAsync(d_scal + i, 0, (size_t)n * sizeof(unsigned long long), stream));
}

unsigned long long MabDev::get_scal(int i)
{
	MAB_CUDA(cudaMemcpyAsync(h_scal, d_scal, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
	MAB_CUDA(cudaStreamSynchronize(stream));
	return h_scal[i];
}

void MabDev::sync()
{
	MAB_CUDA(cudaStreamSynchronize(stream));
}

#include <time.h>
void MabDev::trace(const char *label)
{
	if (trace_on < 0) trace_on = getenv("MAB_TRACE") != nullptr;
	if (!trace_on) return;
	MAB_CUDA(cudaStreamSynchronize(stream));
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	double t = ts.tv_sec + 1e-9 * ts.tv_nsec;
	if (trace_t > 0) fprintf(stderr, "[T]   %-28s %9.3f ms\n", label, (t - trace_t) * 1e3);
	trace_t = t;
}
