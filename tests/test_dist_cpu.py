"""CPU tier, world_size 2 over gloo: the launcher-side logic of the sharded run (byte-range split at line ends,
ownership, hand-over of the communicator id) and the rank aggregation bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from miniasm_b200 import sharded, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, paf, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = open(paf, "rb").read()
    b, e = sharded.split_ranges(data, world)[rank]
    part = data[b:e]
    # every rank's piece ends at a line end; line counts add up
    n_lines = torch.tensor([part.count(b"\n")], dtype=torch.int64)
    dist.all_reduce(n_lines)
    ok = (e == len(data) or data[e - 1:e] == b"\n") and n_lines.item() == data.count(b"\n")
    # the id rank 0 makes reaches everybody unchanged
    box = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ids = [None] * world
    dist.all_gather_object(ids, box[0])
    ok = ok and all(i == ids[0] for i in ids) and len(ids[0]) == 128
    # bench aggregation: throughput = sum of work / max time
    t = torch.tensor([1.0 + rank, 100.0], dtype=torch.float64)
    mx, sm = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ok = ok and mx[0].item() == float(world) and sm[1].item() == 100.0 * world
    q.put((rank, ok, b, e))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_launcher_logic(built, paf_dir):
    paf = synth.generate("-n 1500 -s 77 -j 100", f"{paf_dir}/dist.paf")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, paf, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == os.path.getsize(paf)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_split_ranges_edge_cases(world):
    for data in (b"", b"a\n", b"a\nb", b"\n\n\n", b"x" * 50, b"l1\nl2\nl3\nl4\nl5\nl6\nl7\nl8\nl9\n"):
        rs = sharded.split_ranges(data, world)
        assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == len(data)
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        assert all(e == len(data) or e == b or data[e - 1:e] == b"\n" for b, e in rs)
        assert b"".join(data[b:e] for b, e in rs) == data
    assert [sharded.owner(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]
