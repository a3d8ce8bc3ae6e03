"""One rank of the sharded parity run (launched by tests/test_shard_gpu.py through torch.distributed.run).
argv: paf_path|gen:<workload name of bench.WORKLOADS> out_gfa_path
(gen: = ONE PAF of a BASELINE config cut into `world` parts in file order; every rank generates only its own part)"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miniasm_b200 import capi, sharded  # noqa: E402


def main():
    paf, out = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                      # only carries the 128-byte NCCL id; the data path is NCCL inside the library
    lib = capi.load_product()
    lib.set_verbose(0)
    ctx = lib.mab_create(local)
    sharded.init(lib, ctx, rank, world)
    opt = lib.default_opt()
    stream = os.environ.get("MAB_SHARD_STREAM", "1") != "0"     # load + ingest overlapped (what the CLI and bench.py's e2e leg do) or as two calls
    free = None
    if paf.startswith("gen:"):
        import bench
        buf, n_bytes, _, free = bench.generate_args(bench.WORKLOADS[paf[4:]]["args"], rank, world)
        part, n_part = buf, n_bytes
    else:
        data = open(paf, "rb").read()
        b, e = sharded.split_ranges(data, world)[rank]
        part = data[b:e]
        n_part = len(part)
    if stream:
        lib.mab_load_ingest_text_sharded(ctx, part, n_part, opt.min_span, opt.min_match, 1)
    else:
        lib.mab_load_paf_text(ctx, part, n_part)
        lib.mab_ingest_sharded(ctx, opt.min_span, opt.min_match, 1)
    if free:
        free()                                                  # (both calls synchronise: the host copy can go)
    lib.mab_select_sharded(ctx, C.byref(opt))
    lib.mab_layout_sharded(ctx, C.byref(opt))
    lib.mab_unitigs(ctx)
    if rank == 0:
        st = lib.mab_stats(ctx).contents
        print(f"[shard_worker] world {world}: {st.n_lines} lines on rank 0, {st.n_seq_final} reads kept, {st.n_reduced} arcs reduced, {st.n_utg} unitigs", flush=True)
        if os.environ.get("MAB_GPU_GFA") == "0":           # host structs + the host writer
            d, sub, ug = lib.mab_export_dict(ctx), lib.mab_export_sub(ctx), lib.mab_export_ug(ctx)
            text = lib.print_to_string("ma_ug_print", ug, d, sub)
            with open(out, "wb") as f:
                f.write(text)
        else:                                              # the GFA text formatted on the GPU (what the CLI prints)
            fp = capi._libc.fopen(out.encode(), b"w")
            lib.mab_write_gfa(ctx, fp)
            capi._libc.fclose(fp)
    dist.barrier()
    lib.mab_destroy(ctx)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
