#!/bin/bash
# Round 2, GPU call 4 (1 GPU): GPU tier after the fused ingest / probe-first cleaning / pipelined transitive reduction; bench; ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/r2c4_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|c2 with -f" gpurun_out/r2c4_pytest.log | tail -30
python bench.py --steps 5 --warmup 3 --quick > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c4_bench.json"))
print("value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | launches %s | check %s" % (
    d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['gpu_launches'], d['check']['matches_reference']))
PY
MAB_TRACE=1 python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline 2>&1 >/dev/null | grep -E "^\[T" | tail -45
ncu --set full --clock-control none --import-source on -k regex:"k_del_trans_warp|k_parse" -c 4 -o gpurun_out/r2c4_dt python bench.py --steps 1 --warmup 0 --quick --no-cpu-baseline > gpurun_out/r2c4_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r2c4_dt.ncu-rep --page raw --csv > gpurun_out/r2c4_dt_raw.csv 2>/dev/null
