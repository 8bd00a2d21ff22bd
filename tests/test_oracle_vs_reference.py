"""Pins the CPU oracle (oracle/oracle.cpp) against outputs of the UNMODIFIED reference binary.

Fixtures: tests/golden/<case>.tar.xz, produced by tests/golden/make_golden.py running
oracle/_ref/bin/{plass,penguin} (built from /root/reference by oracle/ref_build.mk).  Every hot-path
step of every iteration is checked independently: the step's golden input DBs go through the oracle
and the result must equal the golden output DB entry by entry (key -> bytes), i.e. bit-exact text.
"""
import os
import numpy as np
import pytest

from common import golden_case, parse_flags, parse_pref_entry
from plass_b200 import mmseqsdb
import oracle_binding as ob
import params

CASES = ["example_aa", "synth_aa", "synth_nt", "long_nt"]


def _steps(case, golden_root, cmd):
    d, man = golden_case(case, golden_root)
    return d, [s for s in man["steps"] if s["cmd"] in cmd]


def hits_from_pref(pref):
    rows = []
    for i, k in enumerate(pref.keys):
        for (t, s, dg) in parse_pref_entry(int(k), pref.entry(i)):
            rows.append((int(k), t, s, dg))
    return np.array(rows, dtype=ob.HIT) if rows else np.zeros(0, dtype=ob.HIT)


def alns_from_db(aln):
    rows = []
    for i, k in enumerate(aln.keys):
        for ln in aln.entry(i).decode().splitlines():
            c = ln.split("\t")
            rows.append((int(k), int(c[0]), int(c[1]), np.float32(float(c[2])), float(c[3]),
                         int(c[4]), int(c[5]), int(c[6]), int(c[7]), int(c[8]), int(c[9])))
    return np.array(rows, dtype=ob.ALN) if rows else np.zeros(0, dtype=ob.ALN)


def assert_same_entries(got, want, what):
    assert set(got) == set(want), what + ": key sets differ"
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, "%s: %d of %d entries differ, first key %d:\n got  %r\n want %r" % (
        what, len(bad), len(want), bad[0], got[bad[0]][:300], want[bad[0]][:300])


@pytest.mark.parametrize("case", CASES)
def test_kmermatcher(case, golden_root):
    d, steps = _steps(case, golden_root, ("kmermatcher",))
    assert steps
    for s in steps:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        want = mmseqsdb.read_db(os.path.join(d, s["dbs"][1]))
        hits = ob.kmermatch(seq, params.oracle_km(s["args"], seq.dbtype == 1))
        got = ob.format_hits_by_rep(seq.keys, hits)
        assert want.dbtype == (14 if seq.dbtype == 1 else 7)
        assert_same_entries(got, want.entries_by_key(), "%s/%s" % (case, s["dbs"][1]))


@pytest.mark.parametrize("case", CASES)
def test_rescorediagonal(case, golden_root):
    d, steps = _steps(case, golden_root, ("rescorediagonal",))
    assert steps
    for s in steps:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        pref = mmseqsdb.read_db(os.path.join(d, s["dbs"][2]))
        want = mmseqsdb.read_db(os.path.join(d, s["dbs"][3]))
        alns = ob.rescore(seq, hits_from_pref(pref), params.oracle_rs(s["args"]))
        got = ob.format_alns_by_query(seq.keys, alns)
        assert_same_entries(got, want.entries_by_key(), "%s/%s" % (case, s["dbs"][3]))


@pytest.mark.parametrize("case", CASES)
def test_assembleresults(case, golden_root):
    d, steps = _steps(case, golden_root, ("assembleresults", "nuclassembleresults"))
    assert steps
    for s in steps:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        aln = mmseqsdb.read_db(os.path.join(d, s["dbs"][1]))
        want = mmseqsdb.read_db(os.path.join(d, s["dbs"][2]))
        alns = alns_from_db(aln)
        # the float parsed from the 3-decimal text must print back to the same text
        assert_same_entries(ob.format_alns_by_query(seq.keys, alns), aln.entries_by_key(), "aln text round trip")
        out, ext = ob.extend(seq, alns, params.oracle_ex(s["args"]))
        assert out.dbtype == want.dbtype
        assert_same_entries(out.entries_by_key(), want.entries_by_key(), "%s/%s" % (case, s["dbs"][2]))
        assert ext.sum() > 0


def test_xxh64_known_answers():
    # SURVEY.md A.1 KATs (vendored xxhash.h, XXH64(le64(v), seed))
    kats = [(67, 0, 0x694b701bc9e44ec7), (67, 1, 0x65d8542382d84f46), (67, 0x0123456789ABCDEF, 0x05ba4c1df800d008),
            (67, 12 ** 14 - 1, 0xae465f8955fd9423), (67, 2 ** 44 - 1, 0xff6d4ee59ddc5537),
            (68, 0, 0xaaa171741b9abdd1), (68, 1, 0x610900b3b71600dc), (68, 0x0123456789ABCDEF, 0x42c4b3605484fb17)]
    for seed, v, h in kats:
        assert ob.lib().or_hash_u64(v, seed) == h


def test_evalue_known_answers():
    import json
    from common import GOLDEN
    t = json.load(open(os.path.join(GOLDEN, "tables.json")))
    L = ob.lib()
    assert L.or_bitscore(0, 255.0) == t["aa_kat_bits255"]
    assert L.or_raw_from_bits(0, 121.0) == t["aa_kat_raw121"]
    assert L.or_evalue(0, 1e8, 255.0, 50.0) == t["aa_kat_eval_255_50_1e8"]
    assert L.or_evalue(0, 1e8, 67.0, 50.0) == t["aa_kat_eval_67_50_1e8"]
    assert L.or_evalue(0, 1e8, 30.0, 21.0) == t["aa_kat_eval_30_21_1e8"]
    assert L.or_bitscore(1, 300.0) == t["nt_kat_bits300"]
    assert L.or_evalue(1, 1.5e8, 300.0, 150.0) == t["nt_kat_eval_300_150_1.5e8"]
    assert L.or_evalue(1, 1.5e8, 98.0, 150.0) == t["nt_kat_eval_98_150_1.5e8"]


# ---- components ranked "next" in SURVEY.md section 8(f): oracle_next.cpp ----------------------------------------------

@pytest.mark.parametrize("case", ["example_aa", "synth_aa"])
def test_findassemblystart(case, golden_root):
    """aa_6f_start_long + aln_0 -> corrected_seqs (data/assemble.sh:108-117), written by the reference binary."""
    d, man = golden_case(case, golden_root)
    seq = mmseqsdb.read_db(os.path.join(d, "aa_6f_start_long"))
    alns = alns_from_db(mmseqsdb.read_db(os.path.join(d, "aln_0")))
    want = mmseqsdb.read_db(os.path.join(d, "corrected_seqs"))
    out, add_stop = ob.findstart(seq, alns)
    assert out.dbtype == want.dbtype == 0
    assert_same_entries(out.entries_by_key(), want.entries_by_key(), "%s/corrected_seqs" % case)
    assert (add_stop >= 0).sum() > 0                      # the fixture does exercise the correction


def test_cyclecheck(golden_root):
    d, man = golden_case("cycle_nt", golden_root)
    seq = mmseqsdb.read_db(os.path.join(d, "seqs"))
    for s in man["steps"]:
        flags = parse_flags(s["args"])
        split = ob.cyclecheck(seq, int(flags["--max-seq-len"]))
        want = mmseqsdb.read_db(os.path.join(d, s["dbs"][1]))
        got = ob.cycle_db(seq, split, int(flags["--chop-cycle"]))
        assert got.n == want.n > 100
        assert_same_entries(got.entries_by_key(), want.entries_by_key(), "cycle_nt/%s" % s["dbs"][1])


@pytest.mark.parametrize("case", ["synth_nt", "long_nt"])
def test_cyclecheck_workflow_split(case, golden_root):
    """nuclassemble.sh:19-60: assembly_N minus the reported sequences = assembly_N_noneCycle (--max-seq-len 200000)."""
    d, man = golden_case(case, golden_root)
    for name in sorted(f for f in os.listdir(d) if f.endswith("_noneCycle")):
        seq = mmseqsdb.read_db(os.path.join(d, name[: -len("_noneCycle")]))
        none = mmseqsdb.read_db(os.path.join(d, name))
        split = ob.cyclecheck(seq, 200000)
        assert sorted(int(k) for k in seq.keys[split == 0]) == sorted(int(k) for k in none.keys), (case, name)


def test_extractorfs_translatenucs(golden_root):
    """nucl_reads -> nucl_<run> (+ header DB) -> aa_<run> for five parameter sets, written by the reference's extractorfs and
    translatenucs --add-orf-stop 1 (reads with N, lower case, IUPAC codes, U, odd lengths, shorter than a codon)."""
    d, man = golden_case("orf_aa", golden_root)
    reads = mmseqsdb.read_db(os.path.join(d, "nucl_reads"))
    runs = [s for s in man["steps"] if s["cmd"] == "extractorfs"]
    assert len(runs) == 5
    # raw reads, --add-orf-stop 0: every length class mod 3, reads shorter than a codon are dropped
    raw = mmseqsdb.read_db(os.path.join(d, "aa_reads"))
    assert raw.n < reads.n
    assert_same_entries(ob.translatenucs(reads).entries_by_key(), raw.entries_by_key(), "orf_aa/aa_reads")
    for s in runs:
        name = s["dbs"][1]
        op = ob.orf_params_from_flags(parse_flags(s["args"]))
        nuc, info = ob.extractorfs(reads, op, False)
        want = mmseqsdb.read_db(os.path.join(d, name))
        assert want.dbtype == 1 and nuc.n == want.n > 200, (name, nuc.n, want.n)
        assert_same_entries(nuc.entries_by_key(), want.entries_by_key(), "orf_aa/" + name)
        assert_same_entries(ob.orf_header_entries(info), mmseqsdb.read_db(os.path.join(d, name + "_h")).entries_by_key(), "orf_aa/%s_h" % name)
        # translatenucs on the reference's own ORF DB, flags from its header DB
        flags = ob.orf_flags_from_headers(mmseqsdb.read_db(os.path.join(d, name + "_h")), want.keys)
        assert_same_entries(ob.translatenucs(want, flags).entries_by_key(),
                            mmseqsdb.read_db(os.path.join(d, "aa_" + name[len("nucl_"):])).entries_by_key(), "orf_aa/translatenucs " + name)
        aa, info2 = ob.extractorfs(reads, op, True)
        assert np.array_equal(info, info2)
        assert_same_entries(aa.entries_by_key(), mmseqsdb.read_db(os.path.join(d, "aa_" + name[len("nucl_"):])).entries_by_key(), "orf_aa/aa_" + name[len("nucl_"):])


def test_translation_against_an_independent_codon_table():
    """The IUPAC state-machine table (TranslateNucl.h) restated in the oracle agrees, on plain ACGT input, with the
    standard genetic code written down independently; ambiguity codes resolve only when every expansion agrees."""
    aa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
    code = {a + b + c: aa[16 * i + 4 * j + k] for i, a in enumerate("TCAG") for j, b in enumerate("TCAG") for k, c in enumerate("TCAG")}
    rng = np.random.default_rng(3)
    seqs = ["".join("ACGT"[x] for x in rng.integers(0, 4, 3 * int(n))) for n in rng.integers(1, 60, 200)]
    db = mmseqsdb.from_sequences([s.encode() for s in seqs], 1)
    out = ob.translatenucs(db).entries_by_key()
    for i, s in enumerate(seqs):
        want = "".join(code[s[j:j + 3]] for j in range(0, len(s), 3))
        assert out[i] == (want + "\n").encode(), (s, out[i])
    # ambiguity: GCN is always Ala, TAR always stop, AAY always Asn, MGR always Arg, NNN unknown, lower case is kept
    amb = mmseqsdb.from_sequences([b"GCNTARAAYMGRNNNgcaGCa"], 1)
    assert ob.translatenucs(amb).entries_by_key()[0] == b"A*NRXaa\n"
