"""Runtime switches of the CUDA library (DESIGN.md "Runtime switches"): every alternative code path must give the
reference's bytes.  The switches are read once per process, so each case runs the command-line tool in a fresh
process with the variable set and compares its GFA with the unmodified reference binary's.

Paths that have not yet been through a GPU run of this suite are only included with MAB_TEST_EXPERIMENTAL=1."""
import os
import subprocess

import pytest

from miniasm_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "miniasm_b200", "miniasm-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "miniasm_ref")

SETS = {
    "chaos": "chaos",
    "bubbles800": "bubbles800",
    "shuffled": "shuffled",
    "deep": "-n 1500 -l 3000 -L 12000 -c 400 -j 30 -s 77",      # several hundred hits per read: the CTA-per-read kernels
    "multi": "-n 5000 -s 5 -d 200000 -j 30",                          # 20 % duplicated overlaps -> multi-arcs (shared marks in the transitive reduction)
    "skew": "skew_small",
    "bubbly": "-n 60000 -l 9000 -L 11000 -j 800 -c 30 -s 21",     # thousands of bubbles/tips along sorted ids: many speculative rounds                                          # hot spots: slabs beyond the warp kernels' limits
}
VERIFIED = [{"MAB_CUB_SELECT": "1"}, {"MAB_SUB_SMEM_SORT": "1"}, {"MAB_WRITER_THREADS": "3"}]
EXPERIMENTAL = [{"MAB_SG_SEGSORT": "1"}, {"MAB_GPU_GFA": "1"}, {"MAB_DT_V7": "1"}, {"MAB_SPEC_WINDOW": "1"}, {"MAB_BUB_EXCUSE": "1"},
                {"MAB_SG_SEGSORT": "1", "MAB_GPU_GFA": "1", "MAB_DT_V7": "1", "MAB_SPEC_WINDOW": "1", "MAB_BUB_EXCUSE": "1"}]


@pytest.fixture(scope="module")
def pafs(built, paf_dir):
    return {k: synth.generate(v, f"{paf_dir}/sw_{k}.paf") for k, v in SETS.items()}


@pytest.fixture(scope="module")
def want(pafs):
    out = {}
    for k, path in pafs.items():
        r = subprocess.run([REF, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0
        out[k] = r.stdout
    return out


def _check(env, name, pafs, want):
    r = subprocess.run([CLI, pafs[name]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env={**os.environ, **env})
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout == want[name]


@pytest.mark.parametrize("name", ["chaos", "bubbles800", "shuffled"])
@pytest.mark.parametrize("env", VERIFIED, ids=lambda e: "+".join(f"{k}={v}" for k, v in e.items()))
def test_switch(env, name, pafs, want):
    _check(env, name, pafs, want)


@pytest.mark.skipif(os.environ.get("MAB_TEST_EXPERIMENTAL") != "1", reason="paths not yet confirmed on a GPU (set MAB_TEST_EXPERIMENTAL=1)")
@pytest.mark.parametrize("name", list(SETS))
@pytest.mark.parametrize("env", EXPERIMENTAL + VERIFIED + [{}], ids=lambda e: "+".join(f"{k}={v}" for k, v in e.items()) or "default")
def test_experimental_switch(env, name, pafs, want):
    _check(env, name, pafs, want)
