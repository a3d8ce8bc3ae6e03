#!/bin/bash
# compute-sanitizer reports with their own log files (stdout of the tool carries the GFA)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, ".")
from miniasm_b200 import synth
synth.generate("chaos_small", "/dev/shm/san.paf")
synth.generate("skew_small", "/dev/shm/san_skew.paf")
PY
CLI=miniasm_b200/miniasm-b200
timeout 300 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 --log-file gpurun_out/r2c9_racecheck.log $CLI /dev/shm/san.paf > /dev/null 2> /dev/null
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r2c9_racecheck.log | sort | uniq -c | sort -rn | head -20
timeout 300 compute-sanitizer --tool memcheck --print-limit 10 --log-file gpurun_out/r2c9_memcheck.log $CLI /dev/shm/san_skew.paf > /dev/null 2> /dev/null
echo "memcheck rc=$?"; tail -2 gpurun_out/r2c9_memcheck.log
timeout 300 compute-sanitizer --tool initcheck --print-limit 10 --log-file gpurun_out/r2c9_initcheck.log $CLI /dev/shm/san.paf > /dev/null 2> /dev/null
echo "initcheck rc=$?"; tail -2 gpurun_out/r2c9_initcheck.log
