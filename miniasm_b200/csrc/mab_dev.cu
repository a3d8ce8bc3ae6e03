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
	cudaMemPool_t pool;
	MAB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
	uint64_t keep = UINT64_MAX; // keep freed blocks cached in the pool: later stages reuse them without driver calls
	MAB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
	MAB_CUDA(cudaMalloc(&d_scal, 64 * sizeof(unsigned long long)));
	MAB_CUDA(cudaMemset(d_scal, 0, 64 * sizeof(unsigned long long)));
	MAB_CUDA(cudaMallocHost(&h_scal, 64 * sizeof(unsigned long long)));
}

void MabDev::destroy()
{
	if (!stream) return;
	MAB_CUDA(cudaSetDevice(device));
	MAB_CUDA(cudaStreamSynchronize(stream));
	if (cub_tmp) MAB_CUDA(cudaFreeAsync(cub_tmp, stream));
	MAB_CUDA(cudaStreamSynchronize(stream));
	MAB_CUDA(cudaFree(d_scal));
	MAB_CUDA(cudaFreeHost(h_scal));
	MAB_CUDA(cudaStreamDestroy(stream));
	stream = nullptr; cub_tmp = nullptr; cub_tmp_bytes = 0; d_scal = h_scal = nullptr;
}

void *MabDev::alloc(size_t bytes)
{
	void *p = nullptr;
	if (bytes == 0) bytes = 16;
	cudaError_t e = cudaMallocAsync(&p, bytes, stream);
	if (e != cudaSuccess) {
		fprintf(stderr, "[E::miniasm_b200] device allocation of %zu bytes failed: %s\n", bytes, cudaGetErrorString(e));
		exit(72);
	}
	return p;
}

void MabDev::free(void *p)
{
	if (p) MAB_CUDA(cudaFreeAsync(p, stream));
}

void *MabDev::tmp(size_t bytes)
{
	if (bytes > cub_tmp_bytes) {
		if (cub_tmp) MAB_CUDA(cudaFreeAsync(cub_tmp, stream));
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
