#!/bin/bash
# Round 2, GPU call 2 (1 GPU): the whole GPU tier at HEAD (fixed-point cleaning passes, new defaults), the bubble-dense
# CLI timing against the reference, the bench line, a launch list and one ncu capture of the transitive reduction.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu =="
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/r2c2_pytest.log
echo "== bubble-dense CLI (600 K reads, jittered ends) vs reference =="
python -c "from miniasm_b200 import synth; synth.generate('-n 600000 -l 9000 -L 11000 -j 800 -c 30 -s 15', '/dev/shm/noisy.paf')"
ls -la /dev/shm/noisy.paf
( time oracle/_ref/miniasm_ref /dev/shm/noisy.paf > /dev/shm/noisy_ref.gfa 2> gpurun_out/r2c2_noisy_ref.err ) 2>&1 | grep real
( time MAB_TRACE=1 miniasm_b200/miniasm-b200 /dev/shm/noisy.paf > /dev/shm/noisy_b200.gfa 2> gpurun_out/r2c2_noisy_b200.err ) 2>&1 | grep real
( time miniasm_b200/miniasm-b200 /dev/shm/noisy.paf > /dev/shm/noisy_b200b.gfa 2> gpurun_out/r2c2_noisy_b200b.err ) 2>&1 | grep real
cmp /dev/shm/noisy_ref.gfa /dev/shm/noisy_b200.gfa && echo "noisy GFA == reference"
grep -E "cleaning passes|popped|cut [0-9]+ tips|T::" gpurun_out/r2c2_noisy_b200.err | head -30
grep -E "popped|Real time" gpurun_out/r2c2_noisy_ref.err | head
echo "== bench (c3) =="
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c2_bench.json"))
print("value %.1f M/s %.2f ms | e2e %.1f M/s %.2f ms | del_trans %.3f ms frac %.3f | phases %s | launches %s" % (
    d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['phase_ms_last_step'], d['gpu_launches']))
print("cpu", d.get("cpu_baseline", {}).get("value"))
PY
echo "== launch list (ncu gpu__time_duration) =="
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2c2_launches.log 2>&1
echo "rc=$?"
echo "== ncu full: transitive reduction + parse =="
ncu --set full --clock-control none --import-source on -k regex:"k_del_trans_warp|k_parse" -c 2 -o gpurun_out/r2c2_dt python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2c2_ncu.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/r2c2_dt.ncu-rep --page raw --csv > gpurun_out/r2c2_dt_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
