#!/bin/bash
# Round 2, 8-GPU call: sharded parity at world 3/4/8 on small sets, the multi-GPU CLI at 8, full-size digests
# (config 3 on 8, config 4 on 4, config 5 skewed on 8), the N=4 / N=8 bench lines, a per-phase trace at N=8.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== full-size digests: c3@8, c4@4, c5@8 =="
T=tests/test_shard_gpu.py
MAB_TEST_FULL=1 timeout 1500 python -m pytest -m gpu -q -s "$T::test_sharded_full_config_digest[c5_8m_skew-8]" "$T::test_sharded_full_config_digest[c3_1m-8]" > gpurun_out/r2m8_full.log 2>&1
echo "rc=$?"; grep -E "passed|failed|shard_worker|Error|error" gpurun_out/r2m8_full.log | tail -12
echo "== small sets at world 3, 4, 8 (2 sets) + CLI at 8 =="
timeout 900 python -m pytest -m gpu -q "$T::test_sharded_gfa_equals_reference[8-chaos]" "$T::test_sharded_gfa_equals_reference[8-bubbles800]" "$T::test_sharded_gfa_equals_reference[4-shuffled]" \
	"$T::test_cli_multi_gpu[ug-8-chaos]" "$T::test_cli_multi_gpu[ug-2-chaos_small]" "$T::test_cli_multi_gpu[sg-3-bubbles800]" > gpurun_out/r2m8_small.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r2m8_small.log
for n in 8; do
echo "== bench N=$n =="
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2m8_bench_n$n.json 2> gpurun_out/r2m8_bench_n$n.err
echo "rc=$?"; python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2m8_bench_n{n}.json").read().strip().splitlines()[-1])
    print("N=%s value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | check %s" % (
        n, d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['check']))
except Exception as e:
    print("bench failed", e)
PY
done
echo "== trace N=8 =="
bash tools/trace_sharded.sh 8 > gpurun_out/r2m8_trace_n8.txt 2>&1; tail -64 gpurun_out/r2m8_trace_n8.txt
