#!/bin/bash
# Round 2, GPU call 5 (1 GPU): streamed load+ingest, the sharded path with one rank (staged push kernel), whole GPU tier, bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest -m gpu -q -x tests/test_hit_gpu.py::test_streamed_ingest_equals_two_calls tests/test_shard_gpu.py::test_sharded_pipeline_world1 "tests/test_configs_gpu.py" > gpurun_out/r2c5_new.log 2>&1
echo "new tests rc=$?"; tail -5 gpurun_out/r2c5_new.log
timeout 1500 python -m pytest tests/test_cli_gpu.py tests/test_switches_gpu.py -m gpu -q --deselect tests/test_cli_gpu.py::test_reads_c2_size > gpurun_out/r2c5_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2c5_pytest.log | tail -20
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c5_bench.json"))
print("value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | launches %s | check %s" % (
    d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['gpu_launches'], d['check']['matches_reference']))
print("cli", d['cli'], "\nnoisy", {k: d['noisy'][k] for k in ('ms_per_step','e2e','cli_wall_s','reference')}, "\nfull", d['cpu_full_size']['seconds'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c5_bench_ref.json 2>/dev/null; echo "ref arm rc=$?"
