#!/bin/bash
# Round 2, GPU call 6 (1 GPU): whole GPU tier at the final code (counting sweep of large ma_hit_sub groups, no-op compaction skipped,
# rank-summed log counts), the bench line, the launch list of one step.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c6_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2c6_pytest.log | tail -20
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c6_bench.json"))
print("value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | launches %s | check %s" % (
    d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['gpu_launches'], d['check']['matches_reference']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2c6_launches.csv python bench.py --steps 1 --warmup 0 --quick --no-cpu-baseline > gpurun_out/r2c6_launches.log 2>&1
echo "launch list rc=$?"
