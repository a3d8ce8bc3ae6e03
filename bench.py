#!/usr/bin/env python
"""bench.py -- PAF overlaps/s, ingest -> GFA, on synthetic PAF of the shapes BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3_1m] [--impl b200|reference]

One "step" = one pass of the whole hot path over one PAF: parse -> read selection -> string graph ->
transitive reduction -> cleaning -> unitigs (main.c:108-199 of the reference).

* value  : whole-job PAF records/s with the PAF bytes already resident in HBM (timed on the device, CUDA
           events on the library's stream, max over ranks).
* e2e    : the same through the C ABI with HOST buffers: pinned PAF text -> H2D -> all steps -> D2H of the
           dictionary / intervals / unitigs -> GFA text written by ma_ug_print to /dev/null (wall clock
           bracketed by device synchronisation; this is the number to hold against the reference arm).
* roofline : asg_arc_del_trans kernel, algorithmic bytes / CUDA-event time vs the measured HBM copy peak.
* cpu_baseline : the unmodified reference (oracle/_ref, single-threaded as it is) on a bounded sample of the
           same workload law, timed on this box's host cores.
--impl reference prints the same line for the reference's own CPU implementation (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from miniasm_b200 import synth  # noqa: E402

# workload -> pafgen options (fixed 10 kb reads, 62.5x, >= 2 kb overlaps, ~50 lines/read)
WORKLOADS = {
    "c2_100k": dict(n_reads=100_000, seed=2, label="Synthetic PAF: 100K reads / 5M overlaps"),
    "c3_1m": dict(n_reads=1_000_000, seed=3, label="Synthetic PAF: 1M reads / 50M overlaps (C. elegans-scale)"),
    "c3_2m": dict(n_reads=2_000_000, seed=4, label="Synthetic PAF: 2M reads / 100M overlaps"),
    "tiny": dict(n_reads=20_000, seed=12, label="Synthetic PAF: 20K reads / 1M overlaps (smoke)"),
}
CPU_SAMPLE_READS = 100_000   # reference arm / cpu_baseline sample: same law, 100K reads ~ 5M PAF lines (~5-8 s of CPU)


class PafgenOpt(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("len_min", C.c_uint32), ("len_max", C.c_uint32), ("coverage", C.c_double),
                ("min_olap", C.c_uint32), ("jitter", C.c_uint32), ("seed", C.c_uint64),
                ("n_hot", C.c_uint32), ("hot_reads", C.c_uint32), ("hot_span", C.c_uint32),
                ("dup_ppm", C.c_uint32), ("self_ppm", C.c_uint32), ("internal_ppm", C.c_uint32), ("lowid_ppm", C.c_uint32),
                ("shuffle", C.c_uint32), ("flip_ppm", C.c_uint32), ("name_base", C.c_uint32)]


class PafgenStat(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("n_bytes", C.c_uint64), ("genome_len", C.c_uint64), ("n_reads_total", C.c_uint32)]


def generate(n_reads, seed, name_base=0):
    """PAF text in C heap memory: (pointer, n_bytes, n_lines, free_fn)."""
    synth.build()
    lib = C.CDLL(synth.LIB)
    lib.pafgen_defaults.argtypes = [C.POINTER(PafgenOpt)]
    lib.pafgen_generate.restype = C.c_size_t
    lib.pafgen_generate.argtypes = [C.POINTER(PafgenOpt), C.POINTER(C.c_void_p), C.POINTER(PafgenStat)]
    lib.pafgen_free.argtypes = [C.c_void_p]
    o, st, buf = PafgenOpt(), PafgenStat(), C.c_void_p()
    lib.pafgen_defaults(C.byref(o))
    o.n_reads, o.seed, o.name_base = n_reads, seed, name_base
    n = lib.pafgen_generate(C.byref(o), C.byref(buf), C.byref(st))
    return buf, n, st.n_lines, lambda: lib.pafgen_free(buf)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.proc = gpu, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        busy = [x for x in sm if x > 0]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(self.rows)}


def run_reference_sample(tmpdir):
    """The unmodified reference (oracle/_ref/miniasm_ref_timed) on the bounded sample; returns its timing dict."""
    exe = os.path.join(ROOT, "oracle", "_ref", "miniasm_ref_timed")
    if not os.path.exists(exe):
        raise FileNotFoundError("oracle/_ref/miniasm_ref_timed missing (run __graft_entry__.build() where /root/reference is mounted)")
    paf = os.path.join(tmpdir, "cpu_sample.paf")
    if not os.path.exists(paf):
        buf, n, n_lines, free = generate(CPU_SAMPLE_READS, 2)
        with open(paf, "wb") as f:
            f.write(C.string_at(buf, n))
        free()
        with open(paf + ".n", "w") as f:
            f.write(str(n_lines))
    n_lines = int(open(paf + ".n").read())
    t0 = time.perf_counter()
    r = subprocess.run([exe, paf], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t0
    tj = [ln for ln in r.stderr.splitlines() if ln.startswith("[T] ")]
    t = json.loads(tj[-1][4:]) if tj else {}
    t["wall"], t["n_lines"] = wall, n_lines
    return t


def host_cpu():
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gpu-gfa", action="store_true", help="e2e leg: GFA text formatted on the GPU (mab_write_gfa, experimental) instead of host structs + ma_ug_print")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[a.workload]
    model, cores = host_cpu()
    tmpdir = tempfile.mkdtemp(prefix="mab_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        secs = []
        t = {}
        for i in range(a.warmup + a.steps):
            t = run_reference_sample(tmpdir)
            if i >= a.warmup:
                secs.append(t["wall"])
        mean = sum(secs) / len(secs)
        v = t["n_lines"] / mean
        sample = f"{CPU_SAMPLE_READS} reads / {t['n_lines']} PAF lines of the same law (fixed 10 kb reads, 62.5x); one full run of the reference per step"
        print(json.dumps({
            "impl": "reference", "metric": "paf_overlaps_per_sec_ingest_to_gfa", "value": v, "unit": "PAF records/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64/uint32 (+3 float32 predicates)", "data": "synthetic",
            "config": {"workload": wl["label"], "sample": sample},
            "cpu_baseline": {"value": v, "unit": "PAF records/s", "cores": 1, "kind": "reference", "sample": sample,
                             "host": f"{model} ({cores} logical cores; the reference is single-threaded)",
                             "del_trans_arcs_per_sec": t.get("n_arc_del_trans_in", 0) / t["asg_arc_del_trans"] if t.get("asg_arc_del_trans") else None},
            "e2e": {"value": v, "unit": "PAF records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ B200 arm
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    saved_stdout = os.dup(1)                  # libraries (NCCL banner ...) must not add lines to stdout: the JSON line is the only one
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from miniasm_b200 import capi
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.load_product()
    lib.set_verbose(0)
    ctx = lib.mab_create(local_rank)
    opt = lib.default_opt()
    if world > 1:
        from miniasm_b200 import sharded
        sharded.init(lib, ctx, rank, world)        # NCCL communicator inside the library (id handed over through torch.distributed)

    # Weak scaling: the job's PAF is the concatenation of one partition of the named shape per rank (disjoint read names,
    # rank order = file order); rank r holds byte range r.  N > 1 runs the hash-sharded pipeline: reads are owned by
    # id mod N, hits travel in an NCCL all-to-all, interval tables / flags are all-reduced, arcs all-gathered.
    buf, n_bytes, n_lines, free = generate(wl["n_reads"], wl["seed"] + 1000 * rank, name_base=rank * 100_000_000)
    pinned = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    C.memmove(pinned.data_ptr(), buf, n_bytes)
    free()
    devnull = capi._libc.fopen(b"/dev/null", b"w")

    def device_steps():
        if world > 1:
            lib.mab_ingest_sharded(ctx, opt.min_span, opt.min_match, 1)
            lib.mab_select_sharded(ctx, C.byref(opt))
            lib.mab_layout_sharded(ctx, C.byref(opt))
        else:
            lib.mab_ingest(ctx, opt.min_span, opt.min_match, 1)
            lib.mab_select(ctx, C.byref(opt), 0, 0, 100)
            lib.mab_layout(ctx, C.byref(opt), 100)
        lib.mab_unitigs(ctx)

    def e2e_step():
        lib.mab_load_paf_text(ctx, pinned.data_ptr(), n_bytes)            # H2D
        device_steps()
        if rank != 0:                                                      # the result is replicated; rank 0 writes it
            return 0
        if a.gpu_gfa:                                                      # experimental: GFA text formatted on the GPU, one D2H
            return lib.mab_write_gfa(ctx, devnull)
        d, sub, ug = lib.mab_export_dict(ctx), lib.mab_export_sub(ctx), lib.mab_export_ug(ctx)   # D2H
        lib.ma_ug_print(ug, d, sub, devnull)                               # GFA text (host C writer)
        nb = 0
        g = ug.contents.g.contents
        nb += (g.n_arc_srt & 0x7fffffff) * 16 + (g.n_seq_symm & 0x7fffffff) * 20
        nb += sum(ug.contents.a[i].n for i in range(min(ug.contents.n, 100000))) * 8 + ug.contents.n * 24
        nb += d.contents.n_seq * (8 + 12)
        lib.ma_ug_destroy(ug), capi.c_free(sub), lib.sd_destroy(d)
        return nb

    def barrier():
        lib.mab_sync(ctx)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- value: inputs resident in HBM, device-timed ------------------------------------------------
    lib.mab_load_paf_text(ctx, pinned.data_ptr(), n_bytes)
    lib.mab_count_del_trans_inner(1)          # one untimed pass with the instrumented kernel: inner-loop iterations I
    device_steps()
    inner = lib.mab_stats(ctx).contents.trans_inner   # (N > 1: this rank's share of the vertices)
    lib.mab_count_del_trans_inner(0)
    for _ in range(a.warmup):
        device_steps()
    st = lib.mab_stats(ctx).contents
    launches0, libcalls0 = st.n_kernel_launches, st.n_lib_calls
    e0, e1 = lib.mab_event_create(), lib.mab_event_create()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dt_ms, t_wall0 = [], time.perf_counter()
    for _ in range(a.steps):
        lib.mab_event_record(ctx, e0)
        device_steps()
        lib.mab_event_record(ctx, e1)
        dt_ms.append(lib.mab_event_elapsed_ms(e0, e1))
        lib.mab_sync(ctx)
        st = lib.mab_stats(ctx).contents
        dt_ms[-1] = (dt_ms[-1], st.ms_del_trans_kernel)
    barrier()
    wall_dev = time.perf_counter() - t_wall0
    st = lib.mab_stats(ctx).contents
    launches, libcalls = st.n_kernel_launches - launches0, st.n_lib_calls - libcalls0
    dev_ms = sum(x[0] for x in dt_ms)
    dt_ms_trans = sum(x[1] for x in dt_ms) / len(dt_ms)
    n_arc_in, n_vtx = st.n_arc_trans_in, 2 * st.n_seq_final // world   # per rank: the vertices this rank reduces

    # ---- e2e: host buffers in, host structures + GFA text out --------------------------------------
    for _ in range(min(a.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(a.steps):
        d2h = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.finish()

    tot = torch.tensor([dev_ms, e2e_s, float(n_lines)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, lines_all = mx[0].item(), mx[1].item(), sm[2].item()
    else:
        lines_all = float(n_lines)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst copy)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        alg_bytes = 16 * n_arc_in + 16 * inner + 1 * n_arc_in + 12 * n_vtx     # DESIGN.md "roofline": own slab + explored neighbour prefixes + flag byte + idx/seq
        achieved = alg_bytes / (dt_ms_trans * 1e-3) / 1e9 if dt_ms_trans else 0.0
        traffic = None
        prof = os.path.join(ROOT, "profiles", "del_trans_r01.json")
        if os.path.exists(prof):
            try:
                pj = json.load(open(prof))
                if pj.get("workload") == a.workload:
                    traffic = pj.get("dram_bytes_per_launch")
            except (OSError, ValueError):
                pass
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            try:
                t = run_reference_sample(tmpdir)
                cpu = {"value": t["n_lines"] / t["total"], "unit": "PAF records/s", "cores": 1, "kind": "reference",
                       "sample": f"{CPU_SAMPLE_READS} reads / {t['n_lines']} PAF lines of the same law, one run of oracle/_ref/miniasm_ref_timed ({t['total']:.2f} s)",
                       "host": f"{model} ({cores} logical cores; the reference is single-threaded)",
                       "del_trans_arcs_per_sec": t["n_arc_del_trans_in"] / t["asg_arc_del_trans"],
                       "seconds_by_function": {k: round(v, 4) for k, v in t.items() if isinstance(v, float) and k not in ("total", "wall")}}
            except Exception as ex:  # noqa: BLE001
                cpu = {"value": None, "error": str(ex)}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps({
            "metric": "paf_overlaps_per_sec_ingest_to_gfa", "value": lines_all * a.steps / (dev_ms * 1e-3), "unit": "PAF records/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64/uint32 (+3 float32 predicates)", "data": "synthetic",
            "config": {"workload": wl["label"], "name": a.workload, "paf_lines_per_gpu": n_lines, "paf_bytes_per_gpu": n_bytes,
                       "l2": "inputs larger than L2 (PAF text and hit arrays are GBs; no flush needed)",
                       "parallelism": (f"read ids hash-sharded over {world} GPUs (owner = id mod {world}); NCCL all-to-all of hits, all-reduce of "
                                       f"interval/flag tables, all-gather of names and surviving arcs, neighbour slabs read from peers over NVLink; one PAF = {world} partitions of the named shape")
                       if world > 1 else "1 GPU"},
            "e2e": {"value": lines_all * a.steps / e2e_s, "unit": "PAF records/s", "ms_per_step": e2e_s / a.steps * 1e3,
                    "h2d_bytes_per_step": n_bytes, "d2h_bytes_per_step": d2h,
                    "writer": "mab_write_gfa (text formatted on the GPU)" if a.gpu_gfa else "mab_export_* + ma_ug_print (host)"},
            "gpu_launches": launches, "lib_calls": libcalls,
            "arcs_per_sec_del_trans": n_arc_in / (dt_ms_trans * 1e-3) if dt_ms_trans else None,
            "del_trans": {"n_arc_in": n_arc_in, "inner_iters": inner, "n_vtx": n_vtx, "kernel_ms": dt_ms_trans},
            "phase_ms_last_step": {"ingest": st.ms_ingest, "select": st.ms_select, "layout": st.ms_layout, "unitigs": st.ms_unitigs},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "k_del_trans_warp",
                         "algorithmic_bytes": alg_bytes, "formula": "16*n_arc + 16*inner_iters + 1*n_arc + 12*n_vtx"},
            "cpu_baseline": cpu, "clocks": clocks, "wall_s_timed_region": wall_dev, "wall_ms_per_step": wall_dev / a.steps * 1e3,
        }), flush=True)
        os.dup2(2, 1)
    lib.mab_event_destroy(e0), lib.mab_event_destroy(e1)
    lib.mab_destroy(ctx)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
