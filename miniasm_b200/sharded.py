"""Launcher-side helpers of the hash-sharded multi-GPU run.

The data path lives in the library (NCCL on the context's stream, csrc/shard_comm.cuh); Python only
(1) cuts the PAF into per-rank byte ranges at line ends, (2) hands the NCCL unique id from rank 0 to the
other ranks through ``torch.distributed`` (any backend), (3) calls the three sharded steps.
"""
import ctypes as C


def split_ranges(data, world):
    """Byte ranges [(begin, end)] of `data` (bytes-like), one per rank, in rank order, each ending after a newline
    (the last one at the end of the data).  Ranges may be empty when there are fewer lines than ranks."""
    n = len(data)
    cuts = [0]
    for r in range(1, world):
        p = max(cuts[-1], n * r // world)
        if p < n:
            q = data.find(b"\n", p)
            p = n if q < 0 else q + 1
        cuts.append(min(p, n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def owner(read_id, world):
    """Rank that owns read `read_id` (ShardComm::owner)."""
    return read_id % world


def exchange_unique_id(lib, rank, world):
    """128 bytes of NCCL unique id, created on rank 0 and broadcast with torch.distributed (needs an initialised group)."""
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if rank == 0:
        lib.mab_nccl_unique_id(buf)
    box = [bytes(buf.raw)]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def init(lib, ctx, rank, world):
    uid = exchange_unique_id(lib, rank, world)
    lib.mab_shard_init(ctx, rank, world, uid)


def run(lib, ctx, opt, bi_dir=1):
    """parse .. reduced + cleaned graph .. unitigs, sharded; every rank ends with the same unitigs."""
    lib.mab_ingest_sharded(ctx, opt.min_span, opt.min_match, bi_dir)
    lib.mab_select_sharded(ctx, C.byref(opt))
    lib.mab_layout_sharded(ctx, C.byref(opt))
    lib.mab_unitigs(ctx)
