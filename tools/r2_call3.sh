#!/bin/bash
# Round 2, GPU call 3 (1 GPU): the whole GPU tier (no -x: every failure in one go) with -R and -f on the GPU.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r2c3_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|c2 with -f" gpurun_out/r2c3_pytest.log | tail -40
