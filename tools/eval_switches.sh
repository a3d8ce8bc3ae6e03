#!/bin/bash
# One GPU call that evaluates every experimental switch: parity first, then A/B bench lines, then an
# ncu capture of the two transitive-reduction kernels.  Usage (from the repo root, under gpurun):
#   gpurun --timeout 1500 -- 'bash tools/eval_switches.sh'            (FULL=1 bash tools/... adds the whole GPU tier under the switches)
# Results land in gpurun_out/sw_*.  Nothing printed by a run under ncu is a bench value.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity of the experimental paths =="
MAB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_switches_gpu.py -m gpu -q > gpurun_out/sw_parity.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/sw_parity.log
run_bench() { # name, env..., -- extra bench args
	local name=$1; shift
	local envs=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
	[ $# -gt 0 ] && shift
	env "${envs[@]}" python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/sw_bench_$name.json 2> gpurun_out/sw_bench_$name.err
	python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/sw_bench_{n}.json"))
    print(f"{n:12s} value {d['value']/1e6:8.1f} M/s  {d['ms_per_step']:6.2f} ms | e2e {d['e2e']['value']/1e6:7.1f} M/s {d['e2e']['ms_per_step']:7.2f} ms | "
          f"del_trans {d['del_trans']['kernel_ms']:.3f} ms frac {d['roofline']['frac']:.3f} | phases {d['phase_ms_last_step']}")
except Exception as e:
    print(n, "FAILED", e)
PY
}
if [ "${FULL:-0}" = "1" ]; then
	# the whole GPU tier with every experimental kernel switched on (the library reads the variables in-process): what has to
	# pass before a switch becomes the default.  ~3.5 min on one GPU.
	echo "== full -m gpu suite with the experimental kernels on =="
	MAB_SG_SEGSORT=1 MAB_DT_V7=1 MAB_BUB_EXCUSE=1 MAB_SPEC_WINDOW=1 MAB_GPU_GFA=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/sw_full_suite.log 2>&1
	echo "rc=$?"; tail -4 gpurun_out/sw_full_suite.log
fi
echo "== bench A/B (1 M reads / 50 M lines) =="
run_bench default
run_bench segsort MAB_SG_SEGSORT=1
run_bench dtv7 MAB_DT_V7=1
run_bench gpugfa -- --gpu-gfa
run_bench all MAB_SG_SEGSORT=1 MAB_DT_V7=1 -- --gpu-gfa
echo "== stage (iii) rounds on a bubble-dense set (300 K reads, jittered ends): default vs windowed =="
python -c "from miniasm_b200 import synth; synth.generate('-n 300000 -l 9000 -L 11000 -j 800 -c 30 -s 15', '/tmp/bub.paf')"
for w in 00 10 01 11; do
	/usr/bin/time -f "window=${w:0:1} excuse=${w:1:1} wall %e s" env MAB_SPEC_WINDOW=${w:0:1} MAB_BUB_EXCUSE=${w:1:1} MAB_TRACE=1 miniasm_b200/miniasm-b200 /tmp/bub.paf > /tmp/bub_$w.gfa 2> gpurun_out/sw_bub_$w.err
	grep -E "cleaning passes|popped|wall" gpurun_out/sw_bub_$w.err | tail -4
	cmp -s /tmp/bub_00.gfa /tmp/bub_$w.gfa && echo "  same GFA as default" || echo "  GFA DIFFERS from default"
done
oracle/_ref/miniasm_ref /tmp/bub.paf 2>/dev/null | cmp -s - /tmp/bub_00.gfa && echo "default GFA == reference" || echo "default GFA differs from the reference"
echo "== ncu: default and v7 transitive reduction =="
for v in 0 1; do
	MAB_DT_V7=$v ncu --set full --clock-control none --import-source on -k regex:"k_del_trans_warp" -c 1 -o gpurun_out/sw_dt_v7_$v \
		python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/sw_ncu_$v.log 2>&1
	ncu -i gpurun_out/sw_dt_v7_$v.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
for row in r[2:]:
    g=lambda k: row[h.index(k)] if k in h else '-'
    print('v7=$v', g('Kernel Name')[:40], 'ms', g('gpu__time_duration.sum'), 'inst', g('smsp__inst_executed.sum'), 'issue%', g('smsp__issue_active.avg.pct_of_peak_sustained_active'), 'L1%', g('l1tex__throughput.avg.pct_of_peak_sustained_active'), 'dramR', g('dram__bytes_read.sum'))
"
done
