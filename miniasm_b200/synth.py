"""Synthetic PAF inputs: thin wrapper over synth/pafgen.c.

``CONFIGS`` names the workloads of BASELINE.json plus the noisy parity variants; every entry is a
pafgen command line, so a PAF is reproducible from (name, pafgen.c) alone.
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "synth", "pafgen.c")
BIN = os.path.join(HERE, "synth", "pafgen")
LIB = os.path.join(HERE, "synth", "libpafgen.so")

CONFIGS = {
    # BASELINE.json configs (exact layout: fixed 10 kb reads, 62.5x, >= 2 kb overlaps, ~50 lines/read)
    "c2_100k": "-n 100000 -s 2",
    "c3_1m": "-n 1000000 -s 3",
    "c3_2m": "-n 2000000 -s 4",          # config 3's law at 2 M reads: the 2-GPU bench workload
    "c4_4m": "-n 4000000 -s 4",
    # config 5: 8 M reads / 400 M overlaps, skewed: 7.98 M reads at 46.9x (~300 M lines) + 2 hot loci of 10 000 reads each
    # (2 x 50 M pairwise lines; those reads have ~10 000 hits and slabs of ~5 000 arcs per vertex)
    "c5_8m_skew": "-n 7980000 -c 46.9 -s 5 -H 2 -R 10000 -W 8000",
    # E. coli-shaped stand-in for config 1 (4.6 Mb, ~30x, variable read length, noisy ends)
    "c1_ecoli_like": "-n 13800 -l 4000 -L 16000 -c 30 -j 200 -s 1",
    # small parity sets: exact, jittered, containment-heavy, tips+bubbles, everything at once
    "tiny_exact": "-n 2000 -s 11",
    "small_exact": "-n 20000 -s 12",
    "jitter30": "-n 20000 -j 30 -s 13",
    "varlen300": "-n 20000 -l 6000 -L 14000 -j 300 -s 14",
    "bubbles800": "-n 30000 -l 9000 -L 11000 -j 800 -c 30 -s 15",
    "chaos": "-n 40000 -l 5000 -L 15000 -j 1200 -c 40 -s 16 -d 2000 -S 500 -I 5000 -D 3000 -C 1000",
    "chaos_small": "-n 6000 -l 5000 -L 15000 -j 1200 -c 40 -s 17 -d 3000 -S 1000 -I 8000 -D 3000 -C 1000",
    "shuffled": "-n 8000 -l 8000 -L 12000 -j 500 -c 35 -s 18 -x",
    "skew_small": "-n 20000 -s 19 -H 4 -R 600 -W 6000",
    "lowcov": "-n 5000 -c 6 -j 100 -s 20",
}


def build():
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-o", BIN, SRC])
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-DPAFGEN_LIB", "-o", LIB, SRC])
    return BIN


def generate(args, out_path):
    """Run pafgen with a command-line string (or a CONFIGS key); returns the output path."""
    build()
    args = CONFIGS.get(args, args)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "wb") as f:
        subprocess.check_call([BIN] + args.split(), stdout=f, stderr=subprocess.DEVNULL)
    return out_path


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def generate_to_file_fast(args, out_path):
    """Same as generate(); kept separate so callers can time it."""
    return generate(args, out_path)
