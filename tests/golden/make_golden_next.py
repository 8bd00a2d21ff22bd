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


if __name__ == "__main__" and "orf" not in sys.argv[1:]:
    main()


# ---- orf_aa: extractorfs + translatenucs ---------------------------------------------------
ORF_RUNS = {
    # the two parameter sets of data/assemble.sh (Assembler.cpp:117-133)
    "start": ["--min-length", "20", "--max-length", "45", "--max-gaps", "0", "--contig-start-mode", "1", "--contig-end-mode", "0", "--orf-start-mode", "0"],
    "long": ["--min-length", "45", "--max-length", "32734", "--max-gaps", "0", "--contig-start-mode", "2", "--contig-end-mode", "2", "--orf-start-mode", "0"],
    # other modes of the ORF finder
    "any": ["--min-length", "10", "--max-length", "32734", "--max-gaps", "3", "--contig-start-mode", "2", "--contig-end-mode", "2", "--orf-start-mode", "1"],
    "last": ["--min-length", "15", "--max-length", "60", "--max-gaps", "1", "--contig-start-mode", "2", "--contig-end-mode", "1", "--orf-start-mode", "2",
             "--use-all-table-starts", "1"],
    "fwd13": ["--min-length", "30", "--max-length", "32734", "--max-gaps", "2147483647", "--contig-start-mode", "0", "--contig-end-mode", "2", "--orf-start-mode", "0",
              "--forward-frames", "1,3", "--reverse-frames", "2"],
}
ORF_COMMON = ["--translation-table", "1", "--translate", "0", "--id-offset", "0", "--create-lookup", "0", "--threads", "1", "--compressed", "0", "-v", "3"]


def orf_reads(seed=41):
    from plass_b200 import synth
    rng = np.random.default_rng(seed)
    reads = [r for r in synth.make_reads(1200, seed=seed)]
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for i, r in enumerate(reads):
        s = np.frombuffer(r if isinstance(r, (bytes, bytearray)) else bytes(r), dtype=np.uint8).copy()
        kind = i % 12
        if kind == 0:                                    # a few N
            p = rng.integers(0, len(s), 3); s[p] = ord("N")
        elif kind == 1:                                  # lower case stretch
            a = int(rng.integers(0, len(s) - 30)); s[a:a + 30] = np.frombuffer(s[a:a + 30].tobytes().lower(), dtype=np.uint8)
        elif kind == 2:                                  # IUPAC ambiguity codes and U / u
            for ch in b"RYKMSWBDHVUu":
                s[int(rng.integers(0, len(s)))] = ch
        elif kind == 3:                                  # length not a multiple of three
            s = s[: len(s) - int(rng.integers(1, 3))]
        elif kind == 4:                                  # merged pair: ~300 nt
            s = np.concatenate([s, acgt[rng.integers(0, 4, int(rng.integers(100, 180)))]])
        elif kind == 5 and i % 24 == 5:                  # very short
            s = s[: int(rng.integers(1, 12))]
        elif kind == 6 and i % 36 == 6:                  # other characters
            s[int(rng.integers(0, len(s)))] = ord("-"); s[int(rng.integers(0, len(s)))] = ord("X")
        out.append(s.tobytes())
    return out


def make_orf_case():
    plass = os.path.join(ROOT, "oracle", "_ref", "bin", "plass")
    work = tempfile.mkdtemp(prefix="golden_orf")
    try:
        seqs = orf_reads()
        src = os.path.join(work, "nucl_reads")
        mmseqsdb.write_db(src, np.arange(len(seqs), dtype=np.uint32), [s + b"\n" for s in seqs], mmseqsdb.DBTYPE_NUCLEOTIDES)
        # extractorfs reads the header DB of its input (only to parse an accession it never uses)
        mmseqsdb.write_db(src + "_h", np.arange(len(seqs), dtype=np.uint32), [b"r%d\n" % i for i in range(len(seqs))], 12)
        pack = os.path.join(work, "pack")
        os.makedirs(pack)
        mmseqsdb.canonicalize(src, os.path.join(pack, "nucl_reads"))
        steps = []
        for name, args in ORF_RUNS.items():
            nuc, aa = os.path.join(work, "nucl_" + name), os.path.join(work, "aa_" + name)
            subprocess.run([plass, "extractorfs", src, nuc] + args + ORF_COMMON, check=True, stdout=subprocess.DEVNULL)
            subprocess.run([plass, "translatenucs", nuc, aa, "--translation-table", "1", "--add-orf-stop", "1", "-v", "3", "--compressed", "0", "--threads", "1"],
                           check=True, stdout=subprocess.DEVNULL)
            for base in ("nucl_" + name, "nucl_" + name + "_h", "aa_" + name):
                mmseqsdb.canonicalize(os.path.join(work, base), os.path.join(pack, base))
            steps.append(dict(cmd="extractorfs", dbs=["nucl_reads", "nucl_" + name], args=args + ORF_COMMON))
            steps.append(dict(cmd="translatenucs", dbs=["nucl_" + name, "aa_" + name], args=["--translation-table", "1", "--add-orf-stop", "1"]))
            print("orf_aa/%s: %d fragments" % (name, mmseqsdb.read_db(os.path.join(pack, "aa_" + name)).n))
        # translatenucs on the raw reads (lengths of every residue class mod 3, reads shorter than a codon), no ORF stops
        aa = os.path.join(work, "aa_reads")
        subprocess.run([plass, "translatenucs", src, aa, "--translation-table", "1", "--add-orf-stop", "0", "-v", "3", "--compressed", "0", "--threads", "1"],
                       check=True, stdout=subprocess.DEVNULL)
        mmseqsdb.canonicalize(aa, os.path.join(pack, "aa_reads"))
        steps.append(dict(cmd="translatenucs", dbs=["nucl_reads", "aa_reads"], args=["--translation-table", "1", "--add-orf-stop", "0"]))
        print("orf_aa/aa_reads: %d of %d reads translated" % (mmseqsdb.read_db(os.path.join(pack, "aa_reads")).n, len(seqs)))
        with open(os.path.join(HERE, "orf_aa.json"), "w") as f:
            json.dump(dict(case="orf_aa", command="plass extractorfs / translatenucs", steps=steps), f, indent=1)
        with tarfile.open(os.path.join(HERE, "orf_aa.tar.xz"), "w:xz") as tf:
            for fn in sorted(os.listdir(pack)):
                tf.add(os.path.join(pack, fn), arcname=fn)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__" and "orf" in sys.argv[1:]:
    make_orf_case()
