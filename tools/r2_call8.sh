#!/bin/bash
# Round 2, last GPU call (1 GPU): smoke(), compute-sanitizer memcheck + racecheck over the command line (default, -R, -f reads).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import sys
sys.path.insert(0, ".")
from miniasm_b200 import synth
paf = synth.generate("chaos_small", "/dev/shm/san.paf")
sys.path.insert(0, "tests")
import importlib
t = importlib.import_module("tests.test_cli_gpu")
t._reads_file(paf, "/dev/shm/san_reads.fa", "fa_var_gaps_extra_dups_missing")
t._reads_file(paf, "/dev/shm/san_reads.fq", "fq_crlf_extra_dups")
synth.generate("-n 4000 -l 2000 -L 30000 -c 30 -j 100 -s 41", "/dev/shm/san_R.paf")
synth.generate("skew_small", "/dev/shm/san_skew.paf")
PY
CLI=miniasm_b200/miniasm-b200
run() { # name, tool, args...
	local name=$1 tool=$2; shift 2
	timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 $CLI "$@" > /dev/shm/san_$name.gfa 2> gpurun_out/r2c8_san_$name.log
	echo "$name ($tool): rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2c8_san_$name.log | tail -1)"
}
run default memcheck /dev/shm/san.paf
run R memcheck -R /dev/shm/san_R.paf
run fa memcheck -f /dev/shm/san_reads.fa /dev/shm/san.paf
run fq memcheck -f /dev/shm/san_reads.fq /dev/shm/san.paf
run skew memcheck /dev/shm/san_skew.paf
run race racecheck /dev/shm/san.paf
for n in default fa; do oracle/_ref/miniasm_ref $( [ $n = fa ] && echo "-f /dev/shm/san_reads.fa" ) /dev/shm/san.paf 2>/dev/null | cmp -s - /dev/shm/san_$n.gfa && echo "$n: GFA under the sanitizer == reference"; done
