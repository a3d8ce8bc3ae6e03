#!/bin/bash
# Round 2, 4-GPU call: sharded parity at world 2/3/4 (fused emit+push exchange and the NCCL route), multi-GPU CLI, config 4 on 4 GPUs,
# bench N=4 (and N=2), per-phase trace at N=4.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
T=tests/test_shard_gpu.py
echo "== small sets, worlds 2/3/4, push route =="
timeout 1200 python -m pytest -m gpu -q "$T::test_sharded_gfa_equals_reference[2-chaos]" "$T::test_sharded_gfa_equals_reference[2-tiny_exact]" "$T::test_sharded_gfa_equals_reference[3-chaos]" \
	"$T::test_sharded_gfa_equals_reference[3-lowcov]" "$T::test_sharded_gfa_equals_reference[4-chaos]" "$T::test_sharded_gfa_equals_reference[4-bubbles800]" "$T::test_sharded_gfa_equals_reference[4-shuffled]" \
	"$T::test_cli_multi_gpu[ug-2-chaos_small]" "$T::test_cli_multi_gpu[ug-4-chaos]" "$T::test_cli_multi_gpu[sg-3-bubbles800]" "$T::test_cli_multi_gpu[ug-3-tiny_exact]" > gpurun_out/r2m4_small.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/r2m4_small.log
echo "== NCCL route (MAB_SHARD_P2P=0) =="
MAB_SHARD_P2P=0 timeout 600 python -m pytest -m gpu -q "$T::test_sharded_gfa_equals_reference[4-chaos]" "$T::test_sharded_gfa_equals_reference[3-shuffled]" > gpurun_out/r2m4_nop2p.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2m4_nop2p.log
echo "== config 4 on 4 GPUs, config 3 on 2 =="
MAB_TEST_FULL=1 timeout 1500 python -m pytest -m gpu -q -s "$T::test_sharded_full_config_digest[c4_4m-4]" "$T::test_sharded_full_config_digest[c3_1m-2]" > gpurun_out/r2m4_full.log 2>&1
echo "rc=$?"; grep -E "passed|failed|shard_worker" gpurun_out/r2m4_full.log | tail -6
for n in 4 2; do
echo "== bench N=$n =="
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2m4_bench_n$n.json 2> gpurun_out/r2m4_bench_n$n.err
echo "rc=$?"; python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2m4_bench_n{n}.json").read().strip().splitlines()[-1])
    print("N=%s value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | check %s" % (
        n, d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['check']['matches_reference']))
except Exception as e:
    print("bench failed", e)
PY
done
echo "== trace N=4 =="
bash tools/trace_sharded.sh 4 > gpurun_out/r2m4_trace_n4.txt 2>&1; tail -64 gpurun_out/r2m4_trace_n4.txt
