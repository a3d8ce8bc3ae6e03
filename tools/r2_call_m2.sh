#!/bin/bash
# Round 2, 2-GPU call: sharded parity at HEAD (world 2, peer reads and all-gather fallback), the multi-GPU CLI, the full-size
# digests of configs 3 and 4 on 2 GPUs, the N=2 bench line and a per-phase trace.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== sharded parity (world 2) + multi-GPU CLI =="
timeout 1500 python -m pytest tests/test_shard_gpu.py -m gpu -x -q > gpurun_out/r2m2_shard.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/r2m2_shard.log
echo "== all-gather fallback (MAB_SHARD_P2P=0) =="
MAB_SHARD_P2P=0 timeout 900 python -m pytest tests/test_shard_gpu.py -m gpu -x -q -k "test_sharded_gfa_equals_reference" > gpurun_out/r2m2_shard_nop2p.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2m2_shard_nop2p.log
echo "== full-size digests: config 3 and config 4 on 2 GPUs =="
MAB_TEST_FULL=1 timeout 1500 python -m pytest tests/test_shard_gpu.py -m gpu -x -q -k "test_sharded_full_config_digest" > gpurun_out/r2m2_full.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/r2m2_full.log
echo "== bench N=2 =="
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2m2_bench_n2.json 2> gpurun_out/r2m2_bench_n2.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2m2_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | check %s" % (
        d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['check']))
except Exception as e:
    print("bench N=2 failed", e)
PY
echo "== trace N=2 =="
bash tools/trace_sharded.sh 2 > gpurun_out/r2m2_trace_n2.txt 2>&1; tail -70 gpurun_out/r2m2_trace_n2.txt
