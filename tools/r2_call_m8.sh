#!/bin/bash
# Round 2, 8-GPU call: the N=8 bench line on config 5 (8 M reads / 401 M lines, skewed; its `check` compares the GFA digest with the
# reference's), a per-phase trace, the multi-GPU command line and one small parity set at world 8.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
T=tests/test_shard_gpu.py
echo "== bench N=8 (config 5) =="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2m8_bench_n8.json 2> gpurun_out/r2m8_bench_n8.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2m8_bench_n8.json").read().strip().splitlines()[-1])
    print("N=8 value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | check %s" % (
        d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['check']))
except Exception as e:
    print("bench failed", e)
PY
tail -5 gpurun_out/r2m8_bench_n8.err
echo "== small set at world 8 + CLI at 8 and 3 =="
timeout 600 python -m pytest -m gpu -q "$T::test_sharded_gfa_equals_reference[8-chaos]" "$T::test_cli_multi_gpu[ug-8-chaos]" "$T::test_cli_multi_gpu[sg-3-bubbles800]" > gpurun_out/r2m8_small.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r2m8_small.log
echo "== trace N=8 =="
timeout 600 bash tools/trace_sharded.sh 8 > gpurun_out/r2m8_trace_n8.txt 2>&1; tail -40 gpurun_out/r2m8_trace_n8.txt
echo "== CLI, config 3 file, 8 GPUs vs 1 =="
python -c "import bench; bench.write_paf(bench.WORKLOADS['c3_1m']['args'], '/dev/shm/c3.paf')"
( time miniasm_b200/miniasm-b200 /dev/shm/c3.paf 2>/dev/null | sha256sum ) 2>&1 | grep -E "real|  -"
( time MINIASM_B200_GPUS=8 miniasm_b200/miniasm-b200 /dev/shm/c3.paf 2>/dev/null | sha256sum ) 2>&1 | grep -E "real|  -"
