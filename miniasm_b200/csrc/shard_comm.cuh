// shard_comm.cuh -- NCCL plumbing of the hash-sharded multi-GPU path: one process per GPU,
// one communicator, every collective issued on the context's stream.  Only the exchanges the path really has:
// all-gather of the distinct read names, all-to-all of hits to the owner of the query read, all-reduce of the
// interval / flag tables, all-gather of arcs.
#pragma once
#include "mab_common.cuh"
#include <nccl.h>
#include <vector>

#define MAB_NCCL(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { \
	fprintf(stderr, "[E::miniasm_b200] %s failed at %s:%d: %s\n", #x, __FILE__, __LINE__, ncclGetErrorString(r_)); exit(79); } } while (0)

struct ShardComm {
	int rank = 0, world = 1;
	ncclComm_t comm = nullptr;
	bool active() const { return world > 1; }
	uint32_t owner(uint32_t read_id) const { return read_id % (uint32_t)world; } // hash-sharding of read ids
};

// sum / max all-reduce in place
static inline void sc_allreduce(MabDev &d, ShardComm &sc, void *buf, size_t count, ncclDataType_t dt, ncclRedOp_t op)
{
	if (sc.active() && count) MAB_NCCL(ncclAllReduce(buf, buf, count, dt, op, sc.comm, d.stream));
}

// one host scalar per rank -> vector of all ranks' values (through a small device buffer)
static inline std::vector<uint64_t> sc_allgather_u64(MabDev &d, ShardComm &sc, uint64_t mine)
{
	std::vector<uint64_t> out((size_t)sc.world, mine);
	if (!sc.active()) return out;
	uint64_t *buf = mab_alloc<uint64_t>(d, (size_t)sc.world + 1);
	MAB_CUDA(cudaMemcpyAsync(buf + sc.world, &mine, 8, cudaMemcpyHostToDevice, d.stream));
	MAB_NCCL(ncclAllGather(buf + sc.world, buf, 1, ncclUint64, sc.comm, d.stream));
	MAB_CUDA(cudaMemcpyAsync(out.data(), buf, 8 * (size_t)sc.world, cudaMemcpyDeviceToHost, d.stream));
	d.sync();
	d.free(buf);
	return out;
}

// variable-size all-gather of bytes: rank r contributes cnt[r] bytes from `mine`; result is the concatenation in rank order
static inline void sc_allgather_v(MabDev &d, ShardComm &sc, const void *mine, const std::vector<uint64_t> &cnt, void *out)
{
	size_t off = 0;
	if (!sc.active()) { if (cnt[0]) MAB_CUDA(cudaMemcpyAsync(out, mine, cnt[0], cudaMemcpyDeviceToDevice, d.stream)); return; }
	MAB_NCCL(ncclGroupStart());
	for (int r = 0; r < sc.world; ++r) {
		if (cnt[r]) MAB_NCCL(ncclBroadcast(r == sc.rank ? mine : (const void*)((char*)out + off), (char*)out + off, cnt[r], ncclUint8, r, sc.comm, d.stream));
		off += cnt[r];
	}
	MAB_NCCL(ncclGroupEnd());
}

// all-to-all of bytes: send_cnt[r] bytes to rank r taken consecutively from `send`; recv_cnt[r] bytes from rank r stored consecutively
static inline void sc_alltoall_v(MabDev &d, ShardComm &sc, const void *send, const std::vector<uint64_t> &send_cnt, void *recv, const std::vector<uint64_t> &recv_cnt)
{
	size_t so = 0, ro = 0;
	MAB_NCCL(ncclGroupStart());
	for (int r = 0; r < sc.world; ++r) {
		if (send_cnt[r]) MAB_NCCL(ncclSend((const char*)send + so, send_cnt[r], ncclUint8, r, sc.comm, d.stream));
		if (recv_cnt[r]) MAB_NCCL(ncclRecv((char*)recv + ro, recv_cnt[r], ncclUint8, r, sc.comm, d.stream));
		so += send_cnt[r], ro += recv_cnt[r];
	}
	MAB_NCCL(ncclGroupEnd());
}
