#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the UNMODIFIED reference binaries
(oracle/_ref/bin/{plass,penguin}, built by oracle/ref_build.mk from /root/reference).

Runs only in the build container (needs /root/reference + oracle/_ref).  For every case it runs the
reference workflow with --remove-tmp-files 0 --delete-tmp-inc 0, then for every hot-path step
(kmermatcher, rescorediagonal, assembleresults / nuclassembleresults) it records the exact argv the
workflow used and stores canonical copies (single data file, key-sorted index) of the step's input
and output DBs.  Result: tests/golden/<case>.tar.xz + tests/golden/<case>.json.

usage: python tests/golden/make_golden.py [case ...]
"""
import json
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from plass_b200 import mmseqsdb  # noqa: E402
from plass_b200 import synth as synth_reads  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
STEPS = ("kmermatcher", "rescorediagonal", "assembleresults", "nuclassembleresults")

CASES = {
    # BASELINE.json configs[0]: the bundled example, 3 iterations so that hash-shift 68 and
    # --include-only-extendable 1 (Assembler.cpp:99-110) are covered as well.
    "example_aa": dict(tool="plass", wf="assemble", iters=3, inputs="example"),
    "synth_aa": dict(tool="plass", wf="assemble", iters=3, inputs=dict(n=4000, seed=11)),
    "synth_nt": dict(tool="penguin", wf="nuclassemble", iters=3, inputs=dict(n=2500, seed=12)),
    # sequences of 32767 residues and more: kmermatcher switches to KmerPosition<int> (kmermatcher.cpp:797-802), diagonals
    # no longer fit 16 bits (hit_t::diagonal wraps, DistanceCalculator.h:94-113 searches the wrapped diagonals)
    "long_nt": dict(tool="penguin", wf="nuclassemble", iters=2, inputs=dict(kind="long_nt", seed=13)),
}


def long_nt_fasta(path, seed):
    """An 80 kb random genome as five long overlapping pieces (45, 50, 70 and 42 kb and a reverse-complemented 40 kb
    one) plus 1500 nearly error-free 150 nt reads on both strands."""
    import numpy as np
    rng = np.random.default_rng(seed)
    g = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 80000)]
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    # g[38000:80000] meets g[0:45000] on diagonal 38000 and g[5000:75000] on 33000: beyond 16 bits
    seqs = [g[0:45000], g[30000:80000], comp[g[20000:60000][::-1]], g[5000:75000], g[38000:80000]]
    for _ in range(1500):
        s = int(rng.integers(0, 80000 - 150))
        r = g[s:s + 150].copy()
        sub = rng.random(150) < 0.003
        r[sub] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(sub.sum()))]
        seqs.append(comp[r[::-1]] if rng.random() < 0.5 else r)
    with open(path, "wb") as f:
        for i, q in enumerate(seqs):
            f.write(b">s%d\n" % i)
            f.write(q.tobytes())
            f.write(b"\n")


def run_case(name, spec, work):
    if spec["inputs"] == "example":
        inputs = ["/root/reference/examples/reads_1.fastq.gz", "/root/reference/examples/reads_2.fastq.gz"]
    elif spec["inputs"].get("kind") == "long_nt":
        fa = os.path.join(work, "reads.fasta")
        long_nt_fasta(fa, spec["inputs"]["seed"])
        inputs = [fa]
    else:
        fa = os.path.join(work, "reads.fasta")
        synth_reads.write_fasta(fa, synth_reads.make_reads(spec["inputs"]["n"], seed=spec["inputs"]["seed"]))
        inputs = [fa]
    tmp = os.path.join(work, "tmp")
    cmd = [os.path.join(REF_BIN, spec["tool"]), spec["wf"]] + inputs + [os.path.join(work, "out.fas"), tmp,
          "--num-iterations", str(spec["iters"]), "--remove-tmp-files", "0", "--delete-tmp-inc", "0", "--threads", "4"]
    log = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
    steps = []
    seen = set()
    for line in log.splitlines():
        w = line.split()
        if not w or w[0] not in STEPS or line in seen:
            continue
        seen.add(line)
        npos = {"kmermatcher": 2, "rescorediagonal": 4}.get(w[0], 3)
        dbs = w[1:1 + npos]
        if not all(d.startswith(tmp) or os.path.realpath(d).startswith(os.path.realpath(tmp)) for d in dbs):
            continue
        steps.append(dict(cmd=w[0], dbs=[os.path.basename(d) for d in dbs], args=w[1 + npos:], paths=dbs))
    out_dir = os.path.join(work, "pack")
    os.makedirs(out_dir)
    done = set()
    for s in steps:
        for base, p in zip(s["dbs"], s["paths"]):
            if base not in done:
                mmseqsdb.canonicalize(p, os.path.join(out_dir, base))
                done.add(base)
        del s["paths"]
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(dict(case=name, command=" ".join(os.path.basename(c) if c.startswith("/") else c for c in cmd), steps=steps), f, indent=1)
    with tarfile.open(os.path.join(HERE, name + ".tar.xz"), "w:xz") as tf:
        for fn in sorted(os.listdir(out_dir)):
            tf.add(os.path.join(out_dir, fn), arcname=fn)
    print(name, "steps:", [(s["cmd"], s["dbs"][-1]) for s in steps])


def main():
    names = sys.argv[1:] or list(CASES)
    for n in names:
        work = tempfile.mkdtemp(prefix="golden_" + n)
        try:
            run_case(n, CASES[n], work)
        finally:
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
