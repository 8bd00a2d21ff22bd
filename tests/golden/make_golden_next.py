#!/usr/bin/env python3
"""Golden fixture for cyclecheck, produced by the UNMODIFIED reference binary
(oracle/_ref/bin/penguin cyclecheck).  Runs only in the build container.

Case cycle_nt: 400 nucleotide sequences -- linear random sequences, circular genomes read past their origin
(terminal redundancy of 5..60 %), tandem repeats, sequences with N runs, sequences shorter than 3 k-mers, lower-case
residues, one sequence above --max-seq-len -- checked with --chop-cycle 0 and 1.
Result: tests/golden/cycle_nt.tar.xz (DBs seqs, cycle_keep, cycle_chop) + cycle_nt.json.
(findassemblystart needs no new fixture: example_aa / synth_aa already hold its input aln_0 and its output corrected_seqs.)
"""
import json
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from plass_b200 import mmseqsdb  # noqa: E402

PENGUIN = os.path.join(ROOT, "oracle", "_ref", "bin", "penguin")
MAX_SEQ_LEN = 20000


def sequences(seed=21):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return acgt[rng.integers(0, 4, n)]

    seqs = []
    for i in range(400):
        kind = i % 8
        if kind in (0, 1):                                  # linear
            s = rnd(int(rng.integers(30, 3000)))
        elif kind in (2, 3, 4):                             # circular genome read past its origin
            g = rnd(int(rng.integers(120, 4000)))
            extra = max(1, int(len(g) * rng.uniform(0.05, 0.6)))
            s = np.concatenate([g, g[:extra]])
            if kind == 4:                                   # with a few substitutions in the redundant part
                m = rng.random(len(s)) < 0.01
                s = s.copy(); s[m] = acgt[rng.integers(0, 4, int(m.sum()))]
        elif kind == 5:                                     # tandem repeat
            u = rnd(int(rng.integers(25, 400)))
            s = np.tile(u, int(rng.integers(2, 6)))[: int(rng.integers(60, 1500))]
        elif kind == 6:                                     # N runs / lower case
            g = rnd(int(rng.integers(200, 2000)))
            s = np.concatenate([g, g[: len(g) // 3]]).copy()
            p = int(rng.integers(0, len(s) - 10))
            s[p:p + int(rng.integers(1, 10))] = ord("N")
            if i % 16 == 6:
                s = np.frombuffer(s.tobytes().lower(), dtype=np.uint8)
        else:                                               # very short
            s = rnd(int(rng.integers(1, 70)))
        seqs.append(s.tobytes())
    g = rnd(15000)
    seqs.append(np.concatenate([g, g[:6000]]).tobytes())    # 21000 >= --max-seq-len: skipped by the reference
    g = rnd(9000)
    seqs.append(np.concatenate([g, g[:4000]]).tobytes())    # a long circular one below the limit
    return seqs


def main():
    work = tempfile.mkdtemp(prefix="golden_cycle")
    try:
        seqs = sequences()
        keys = np.arange(len(seqs), dtype=np.uint32) * 3 + 1          # non-dense keys
        src = os.path.join(work, "seqs")
        mmseqsdb.write_db(src, keys, [s + b"\n" for s in seqs], mmseqsdb.DBTYPE_NUCLEOTIDES)
        pack = os.path.join(work, "pack")
        os.makedirs(pack)
        mmseqsdb.canonicalize(src, os.path.join(pack, "seqs"))
        steps = []
        for name, chop in (("cycle_keep", 0), ("cycle_chop", 1)):
            out = os.path.join(work, name)
            args = ["--max-seq-len", str(MAX_SEQ_LEN), "--chop-cycle", str(chop), "--threads", "4", "-v", "3"]
            subprocess.run([PENGUIN, "cyclecheck", src, out] + args, check=True, stdout=subprocess.DEVNULL)
            mmseqsdb.canonicalize(out, os.path.join(pack, name))
            steps.append(dict(cmd="cyclecheck", dbs=["seqs", name], args=args))
        with open(os.path.join(HERE, "cycle_nt.json"), "w") as f:
            json.dump(dict(case="cycle_nt", command="penguin cyclecheck seqs <out> ...", steps=steps), f, indent=1)
        with tarfile.open(os.path.join(HERE, "cycle_nt.tar.xz"), "w:xz") as tf:
            for fn in sorted(os.listdir(pack)):
                tf.add(os.path.join(pack, fn), arcname=fn)
        db = mmseqsdb.read_db(os.path.join(pack, "cycle_chop"))
        print("cycle_nt: %d sequences, %d reported circular" % (len(seqs), db.n))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
