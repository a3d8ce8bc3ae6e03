#!/bin/bash
# last check of round 2: racecheck / initcheck after the two sanitizer-motivated changes, transitive-reduction parity, a quick bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, ".")
from miniasm_b200 import synth
synth.generate("chaos_small", "/dev/shm/san.paf")
PY
CLI=miniasm_b200/miniasm-b200
timeout 300 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10 --log-file gpurun_out/r2c10_racecheck.log $CLI /dev/shm/san.paf > /dev/shm/san_rc.gfa 2> /dev/null
echo "racecheck rc=$?: $(grep -E 'RACECHECK SUMMARY' gpurun_out/r2c10_racecheck.log)"
timeout 300 compute-sanitizer --tool initcheck --print-limit 10 --log-file gpurun_out/r2c10_initcheck.log $CLI /dev/shm/san.paf > /dev/null 2> /dev/null
echo "initcheck rc=$?: $(tail -1 gpurun_out/r2c10_initcheck.log)"
oracle/_ref/miniasm_ref /dev/shm/san.paf 2>/dev/null | cmp -s - /dev/shm/san_rc.gfa && echo "GFA under racecheck == reference"
timeout 600 python -m pytest tests/test_asg_gpu.py -m gpu -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f M/s %.2f ms | e2e %.2f ms | del_trans %.3f ms frac %.3f | check %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['check']['matches_reference']))"
