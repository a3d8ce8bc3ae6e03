"""CPU tier: the oracle port (oracle/ma_oracle.c) and the host-side writers against the committed golden
vectors that the unmodified reference produced (tests/golden/make_golden.py)."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from miniasm_b200 import capi, synth
from oracle import loaders
from miniasm_b200.pipeline import Pipeline

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = ["tiny_exact", "chaos_small", "lowcov", "shuffled", "jitter30"]


def gold(name):
    return json.load(open(os.path.join(GOLD, name + ".json")))


def sha_sorted(text):
    return hashlib.sha256(b"\n".join(sorted(text.split(b"\n")))).hexdigest()


@pytest.fixture(scope="module")
def port(built):
    lib = loaders.load_oracle_port()
    return lib


@pytest.fixture(scope="module")
def pafs(built, paf_dir):
    return {n: synth.generate(n, f"{paf_dir}/{n}.paf") for n in SETS}


@pytest.mark.parametrize("name", SETS)
def test_generator_is_deterministic(name, pafs):
    g = gold(name)
    assert synth.CONFIGS[name] == g["pafgen_args"]
    assert synth.sha256(pafs[name]) == g["paf_sha256"]


@pytest.mark.parametrize("name", SETS)
def test_oracle_port_reproduces_reference_gfa(name, pafs, port):
    g = gold(name)
    text = Pipeline(port, pafs[name]).run_all()
    assert hashlib.sha256(text).hexdigest() == g["gfa_sha256"]
    path = os.path.join(GOLD, name + ".gfa.gz")
    if os.path.exists(path):
        assert text == gzip.open(path).read()


@pytest.mark.parametrize("name", ["chaos_small", "lowcov", "shuffled"])
@pytest.mark.parametrize("stage", [5, 6, 7, 9, 10, 100])
def test_oracle_port_stage_graphs(name, stage, pafs, port):
    """`-S k -p sg` dumps of the reference (arcs after each cleaning step) as sorted multisets."""
    g = gold(name)
    key = "-p sg" if stage == 100 else f"-S{stage} -p sg"
    p = Pipeline(port, pafs[name]).read().select().sg_gen().clean(upto=stage)
    assert sha_sorted(p.sg_text()) == g["dumps"][key]["sorted_sha256"]
    p.free()


@pytest.mark.parametrize("name", ["chaos_small", "shuffled"])
def test_oracle_port_stage_hits(name, pafs, port):
    """`-S k -p paf` dumps: hits after sub+cut / flt / 2nd sub+cut / containment, rebuilt from the port's arrays
    in the reference's print format (main.c:21-30)."""
    g = gold(name)

    def dump(p):
        h, s, nm = p.hits_np(), p.sub_np(), p.names()
        out = []
        for r in h:
            q, t = int(r["qns"] >> 32), int(r["tn"])
            qs, qe = int(s[q]["s_del"] & 0x7fffffff), int(s[q]["e"])
            ts, te = int(s[t]["s_del"] & 0x7fffffff), int(s[t]["e"])
            out.append("%s:%d-%d\t%d\t%d\t%d\t%s\t%s:%d-%d\t%d\t%d\t%d\t%d\t%d\t255" % (
                nm[q].decode(), qs + 1, qe, qe - qs, int(r["qns"] & 0xffffffff), r["qe"], "+-"[int(r["ml_rev"]) >> 31],
                nm[t].decode(), ts + 1, te, te - ts, r["ts"], r["te"], int(r["ml_rev"]) & 0x7fffffff, int(r["bl_del"]) & 0x7fffffff))
        return ("\n".join(sorted(out + [""]))).encode()

    p = Pipeline(port, pafs[name]).read().sub1().cut()
    assert hashlib.sha256(dump(p)).hexdigest() == g["dumps"]["-S2 -p paf"]["sorted_sha256"]
    p.flt()
    assert hashlib.sha256(dump(p)).hexdigest() == g["dumps"]["-S3 -p paf"]["sorted_sha256"]
    p.sub2_cut_merge()
    assert hashlib.sha256(dump(p)).hexdigest() == g["dumps"]["-S4 -p paf"]["sorted_sha256"]
    p.contained()
    assert hashlib.sha256(dump(p)).hexdigest() == g["dumps"]["-S5 -p paf"]["sorted_sha256"]
    p.free()


def test_host_writers_match_golden(pafs, port, built):
    """The product's host C writers (gfa.c) print the reference's text for the port's structures."""
    prod = capi.load_product()
    name = "chaos_small"
    p = Pipeline(port, pafs[name]).read().select().sg_gen().clean().ug_gen()
    text = prod.print_to_string("ma_ug_print", p.ug, p.d, p.sub)
    assert hashlib.sha256(text).hexdigest() == gold(name)["gfa_sha256"]
    assert sha_sorted(prod.print_to_string("ma_sg_print", p.sg, p.d, p.sub)) == gold(name)["dumps"]["-p sg"]["sorted_sha256"]
    p.free()


@pytest.mark.skipif(not os.path.exists(os.path.join(capi.ROOT, "oracle", "_ref", "miniasm_ref")), reason="reference binary not built")
@pytest.mark.parametrize("name", ["chaos_small", "lowcov"])
def test_golden_files_are_current(name, pafs):
    import subprocess
    out = subprocess.run([os.path.join(capi.ROOT, "oracle", "_ref", "miniasm_ref"), pafs[name]], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    assert hashlib.sha256(out).hexdigest() == gold(name)["gfa_sha256"]


def test_oracle_port_next_rows(port, paf_dir):
    """SURVEY.md 8f rows 2 and 3 against vectors the reference binary produced (tests/golden/next_rows.json, make_golden.py next):
    `-R` (ma_hit_no_cont + ma_hit_read with its exclusion list) and `-f reads` (ma_ug_seq) through the port."""
    import ctypes as C

    from tests.golden.make_golden import NEXT_F
    from tests.test_cli_gpu import _reads_file
    g = json.load(open(os.path.join(GOLD, "next_rows.json")))
    paf = synth.generate(g["R"]["pafgen_args"], f"{paf_dir}/next_R.paf")
    assert synth.sha256(paf) == g["R"]["paf_sha256"]
    p = Pipeline(port, paf)
    o = p.opt
    excl = port.ma_hit_no_cont(p.paf, o.min_span, o.min_match, o.max_hang, o.int_frac)
    assert f"dropped {excl.contents.n_seq} contained reads" == g["R"]["dropped"][0]
    p.d, n = port.sd_init(), C.c_size_t(0)
    p.hits = port.ma_hit_read(p.paf, o.min_span, o.min_match, p.d, C.byref(n), 1, excl)
    p.n_hits = n.value
    assert f"stored {p.n_hits} hits and {p.d.contents.n_seq} sequences" in g["R"]["stderr_counts"][0]
    text = p.select().sg_gen().clean().ug_gen().gfa()
    assert hashlib.sha256(text).hexdigest() == g["R"]["gfa_sha256"]
    p.free(), port.sd_destroy(excl)
    paf = synth.generate("chaos_small", f"{paf_dir}/chaos_small.paf")
    for fn in NEXT_F:
        style, ext = fn.rsplit(".", 1)
        reads = _reads_file(paf, f"{paf_dir}/next_{fn}", style)
        assert hashlib.sha256(open(reads, "rb").read()).hexdigest() == g["f"][fn]["reads_sha256"]
        text = Pipeline(port, paf).run_all(reads)
        assert len(text) == g["f"][fn]["gfa_bytes"] and hashlib.sha256(text).hexdigest() == g["f"][fn]["gfa_sha256"], fn
