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
