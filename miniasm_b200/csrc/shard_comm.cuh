// shard_comm.cuh -- NCCL plumbing of the hash-sharded multi-GPU path: one process per GPU,
// one communicator, every collective issued on the context's stream.  Only the exchanges the path really has:
// all-gather of the distinct read names, all-to-all of hits to the owner of the query read, all-reduce of the
// interval / flag tables, all-gather of arcs.
#pragma once
#include "mab_common.cuh"
#include <nccl.h>
#include <vector>
#include <map>
#include <string>
#include <unistd.h>

#define MAB_NCCL(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { \
	fprintf(stderr, "[E::miniasm_b200] %s failed at %s:%d: %s\n", #x, __FILE__, __LINE__, ncclGetErrorString(r_)); exit(79); } } while (0)

struct ShardComm {
	int rank = 0, world = 1;
	ncclComm_t comm = nullptr;
	std::map<std::string, void*> ipc_open;   // peer segments mapped through CUDA IPC (handle bytes -> local address), kept for the life of the context
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

// Addresses under which this rank can load from / store to a buffer of every rank (NVLink peer access): `mine` is this rank's
// buffer (inside an arena segment), out[r] the address of rank r's.  Ranks in other PROCESSES are reached through CUDA IPC handles
// of their arena segment, ranks that are threads of THIS process through plain peer access (an IPC handle cannot be opened by the
// process that exported it).  Collective: every rank calls it; false (on all ranks alike) if any rank cannot offer or reach a
// buffer, or MAB_SHARD_P2P=0 -- the callers then take their NCCL route.
static inline bool sc_peer_ptrs(MabDev &d, ShardComm &sc, const void *mine_ptr, std::vector<void*> &out)
{
	const int G = sc.world;
	out.assign((size_t)G, nullptr);
	struct PeerInfo { cudaIpcMemHandle_t h; uint64_t off, ok, pid, ptr, dev; };
	PeerInfo mine;
	memset(&mine, 0, sizeof(mine));
	{
		char *base; size_t off;
		const char *env = getenv("MAB_SHARD_P2P");
		mine.pid = (uint64_t)getpid(), mine.ptr = (uint64_t)(uintptr_t)mine_ptr, mine.dev = (uint64_t)d.device;
		if (!(env && atoi(env) == 0) && d.arena.segment_of(mine_ptr, &base, &off) && cudaIpcGetMemHandle(&mine.h, base) == cudaSuccess) mine.off = off, mine.ok = 1;
		else cudaGetLastError();
	}
	std::vector<PeerInfo> peers((size_t)G);
	{
		PeerInfo *buf = (PeerInfo*)d.alloc(sizeof(PeerInfo) * ((size_t)G + 1));
		MAB_CUDA(cudaMemcpyAsync(buf + G, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice, d.stream));
		if (sc.active()) MAB_NCCL(ncclAllGather(buf + G, buf, sizeof(PeerInfo), ncclUint8, sc.comm, d.stream));
		else MAB_CUDA(cudaMemcpyAsync(buf, buf + G, sizeof(PeerInfo), cudaMemcpyDeviceToDevice, d.stream));
		MAB_CUDA(cudaMemcpyAsync(peers.data(), buf, sizeof(PeerInfo) * (size_t)G, cudaMemcpyDeviceToHost, d.stream));
		d.sync();
		d.free(buf);
	}
	bool p2p = true;
	for (int r = 0; r < G && p2p; ++r) {
		if (!peers[r].ok) { p2p = false; break; }
		if (r == sc.rank) { out[r] = (void*)mine_ptr; continue; }
		if (peers[r].pid == mine.pid) { // same process: direct peer access
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, d.device, (int)peers[r].dev) != cudaSuccess || !can) { cudaGetLastError(); p2p = false; break; }
			cudaError_t e = cudaDeviceEnablePeerAccess((int)peers[r].dev, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); p2p = false; break; }
			cudaGetLastError();
			out[r] = (void*)(uintptr_t)peers[r].ptr;
			continue;
		}
		std::string key((const char*)&peers[r].h, sizeof(cudaIpcMemHandle_t));
		auto it = sc.ipc_open.find(key);
		void *base = nullptr;
		if (it != sc.ipc_open.end()) base = it->second;
		else if (cudaIpcOpenMemHandle(&base, peers[r].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) sc.ipc_open[key] = base;
		else { cudaGetLastError(); p2p = false; break; }
		out[r] = (char*)base + peers[r].off;
	}
	std::vector<uint64_t> okv = sc_allgather_u64(d, sc, p2p ? 1 : 0); // all ranks take the same route
	for (int r = 0; r < G; ++r) p2p = p2p && okv[r] != 0;
	return p2p;
}
