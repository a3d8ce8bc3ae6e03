"""Multi-GPU parity: the hash-sharded pipeline on 2 (or more) GPUs must print the GFA the reference
prints for the whole PAF -- byte for byte.  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest

from miniasm_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "miniasm_ref")


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


WORLDS = [w for w in (2, 3, 4, 8) if w <= n_gpus()]


@pytest.mark.skipif(n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("name", ["chaos_small", "chaos", "bubbles800", "tiny_exact", "shuffled", "lowcov"])
@pytest.mark.parametrize("world", WORLDS or [2])
def test_sharded_gfa_equals_reference(name, world, built, paf_dir):
    paf = synth.generate(name, f"{paf_dir}/sh_{name}.paf")
    out = f"{paf_dir}/sh_{name}_{world}.gfa"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "shard_worker.py"), paf, out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    want = subprocess.run([REF, paf], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    assert open(out, "rb").read() == want


@pytest.mark.skipif(n_gpus() < 2 or os.environ.get("MAB_TEST_FULL") != "1", reason="needs >= 2 GPUs and MAB_TEST_FULL=1 (minutes)")
@pytest.mark.parametrize("name,n_reads,seed", [("c3_1m", 1_000_000, 3), ("c4_4m", 4_000_000, 4)])
@pytest.mark.parametrize("world", [w for w in (2, 4, 8) if w <= n_gpus()] or [2])
def test_sharded_full_config_digest(name, n_reads, seed, world, built, paf_dir):
    """BASELINE configs 3 / 4 (1 M / 4 M reads) hash-sharded over `world` GPUs: the GFA must have the sha256 of the
    reference's GFA for the same PAF (tests/golden/configs.json)."""
    import hashlib
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "configs.json")))
    if name not in gold:
        pytest.skip(f"no golden digest for {name}")
    out = f"{paf_dir}/shfull_{name}_{world}.gfa"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tests", "shard_worker.py"), f"gen:{n_reads}:{seed}", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-3000:]
    got = open(out, "rb").read()
    os.unlink(out)
    assert len(got) == gold[name]["gfa_bytes"] and hashlib.sha256(got).hexdigest() == gold[name]["gfa_sha256"]
