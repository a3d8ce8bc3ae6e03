#!/bin/bash
# Round 2, last 2-GPU call: the multi-GPU command line (stdout + rank-summed counters), both load paths of the sharded worker, bench N=2.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=tests/test_shard_gpu.py
timeout 900 python -m pytest -m gpu -q "$T::test_cli_multi_gpu[ug-2-chaos_small]" "$T::test_cli_multi_gpu[sg-2-bubbles800]" "$T::test_cli_multi_gpu[ug-2-chaos]" \
	"$T::test_sharded_gfa_equals_reference[2-chaos]" "$T::test_sharded_gfa_equals_reference[2-shuffled]" > gpurun_out/r2m2b_tests.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/r2m2b_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2m2b_bench_n2.json 2> gpurun_out/r2m2b_bench_n2.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2m2b_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | phases %s | check %s" % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['phase_ms_last_step'], d['check']['matches_reference']))
except Exception as e:
    print("bench failed", e)
PY
