#!/usr/bin/env python
"""Known answers for BASELINE.json's synthetic configurations at FULL size: the unmodified reference binary
(oracle/_ref/miniasm_ref) is run on the generated PAF and the digests of its GFA are written to
tests/golden/configs.json.  The inputs are not stored (GBs): (pafgen options, sha256 of the PAF) identify them, and
the GPU tests regenerate the same bytes in memory (tests/test_configs_gpu.py).

usage: python tests/golden/make_golden_configs.py [c2_100k c3_1m c4_4m]      (c3 needs ~10 GB RAM / 2 min, c4 ~25 GB / 7 min, c5 ~45 GB / 20 min)
"""
import hashlib
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from miniasm_b200 import synth  # noqa: E402
from make_golden import counts, sha_lines  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "miniasm_ref")
OUT = os.path.join(HERE, "configs.json")


def main():
    names = sys.argv[1:] or ["c2_100k", "c3_1m"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        paf = synth.generate(name, f"/tmp/mab_golden_{name}.paf")
        t0 = time.time()
        r = subprocess.run([REF, paf], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
        gfa = r.stdout
        res[name] = {"pafgen_args": synth.CONFIGS[name], "paf_sha256": synth.sha256(paf), "paf_bytes": os.path.getsize(paf),
                     "gfa_bytes": len(gfa), "gfa_sha256": hashlib.sha256(gfa).hexdigest(), "gfa_sorted_sha256": sha_lines(gfa, True),
                     "stderr_counts": counts(r.stderr.decode()), "reference_seconds": round(time.time() - t0, 1)}
        print(name, res[name]["gfa_bytes"], res[name]["gfa_sha256"][:16], res[name]["reference_seconds"], "s", flush=True)
        os.unlink(paf)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
