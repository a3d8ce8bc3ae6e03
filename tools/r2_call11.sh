#!/bin/bash
# very last GPU call of round 2: smoke(), a quick bench line at HEAD, then config 4 (4 M reads / 200 M lines) on ONE GPU against the reference's digest
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f M/s %.2f ms | e2e %.2f ms | del_trans %.3f ms frac %.3f | check %s | cleaning %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step'], d['del_trans']['kernel_ms'], d['roofline']['frac'], d['check']['matches_reference'], d['cleaning_passes']))"
MAB_TEST_FULL=1 timeout 300 python -m pytest -m gpu -q -s "tests/test_configs_gpu.py::test_full_size_config_matches_reference_digest[c4_4m-4000000-4-stream]" 2>&1 | tail -3
