#!/bin/bash
# Phase-by-phase wall times of the hash-sharded pipeline (MAB_TRACE=1: every traced step synchronises, so the lines are
# a breakdown, not a bench value).  Usage: gpurun --gpus N --timeout 900 -- 'bash tools/trace_sharded.sh N'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
MAB_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
	bench.py --gpus "$N" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/trace_n$N.json 2> gpurun_out/trace_n$N.err
echo "rc=$?"
# the last traced pass of rank 0 (lines are interleaved between ranks; rank 0 prints the [M::] lines too)
grep -E "^\[T" gpurun_out/trace_n$N.err | tail -120 | awk '{k=$0; sub(/[0-9.]+ ms.*/,"",k); c[k]++; t[k]+=$(NF-1)} END {for (k in c) printf "%8.3f ms avg over %3d  %s\n", t[k]/c[k], c[k], k}' | sort -rn | head -90
tail -1 gpurun_out/trace_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phase_ms_last_step'])"
