"""GPU parity tests (-m gpu) of the two per-iteration helpers next to the hot path:
findassemblystart and cyclecheck, through the C ABI and through the host CLI, against the CPU oracle (oracle_next.cpp)
and against DBs written by the unmodified reference binary (tests/golden)."""
import os
import subprocess
import numpy as np
import pytest

from common import golden_case, parse_flags, ROOT
from plass_b200 import mmseqsdb, api
import oracle_binding as ob
from test_oracle_vs_reference import alns_from_db, assert_same_entries

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "plass_b200", "plass_b200_cli")


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", ["example_aa", "synth_aa"])
def test_findassemblystart_matches_oracle_and_golden(case, golden_root, ctx):
    d, man = golden_case(case, golden_root)
    seq = mmseqsdb.read_db(os.path.join(d, "aa_6f_start_long"))
    alns = alns_from_db(mmseqsdb.read_db(os.path.join(d, "aln_0")))
    ddb = ctx.upload(seq)
    out, add_stop = ctx.findassemblystart(ddb, alns)
    got = out.download()
    out.free(); ddb.free()
    want, wstop = ob.findstart(seq, alns)
    assert np.array_equal(add_stop, wstop), case
    assert_same_entries(got.entries_by_key(), want.entries_by_key(), "%s/corrected_seqs vs oracle" % case)
    golden = mmseqsdb.read_db(os.path.join(d, "corrected_seqs"))
    assert_same_entries(got.entries_by_key(), golden.entries_by_key(), "%s/corrected_seqs vs reference" % case)
    assert ctx.timings()["kernel_launches"] > 0


@pytest.mark.parametrize("case", ["example_aa", "synth_aa"])
def test_fused_step0_matches_golden_chain(case, golden_root, ctx):
    """pg_assemble_step0: aa_6f_start_long -> corrected_seqs -> pref_corrected_0 / aln_corrected_0 -> assembly_0 in HBM."""
    import params
    from test_gpu_parity import gpu_km
    d, man = golden_case(case, golden_root)
    steps = man["steps"]
    km = [s for s in steps if s["cmd"] == "kmermatcher" and s["dbs"][0] == "aa_6f_start_long"][0]
    rs = [s for s in steps if s["cmd"] == "rescorediagonal" and s["dbs"][0] == "aa_6f_start_long"][0]
    ex = [s for s in steps if s["cmd"] == "assembleresults" and s["dbs"][2] == "assembly_0"][0]
    seq = mmseqsdb.read_db(os.path.join(d, "aa_6f_start_long"))
    ddb = ctx.upload(seq)
    corr, out, hits, alns = ctx.assemble_step0(ddb, gpu_km(km["args"], False), api.RsParams(**params.rs_fields(rs["args"])),
                                               api.ExParams(**params.ex_fields(ex["args"])), want_intermediates=True)
    got_corr, got_out = corr.download(), out.download()
    corr.free(); out.free(); ddb.free()
    assert_same_entries(got_corr.entries_by_key(), mmseqsdb.read_db(os.path.join(d, "corrected_seqs")).entries_by_key(), "%s/corrected_seqs fused" % case)
    pref = mmseqsdb.read_db(os.path.join(d, "pref_corrected_0"))
    assert_same_entries(ob.format_hits_by_rep(got_corr.keys, hits), pref.entries_by_key(), "%s/pref_corrected_0 fused" % case)
    assert_same_entries(got_out.entries_by_key(), mmseqsdb.read_db(os.path.join(d, "assembly_0")).entries_by_key(), "%s/assembly_0 fused" % case)


def test_findassemblystart_without_alignments(ctx):
    seq = mmseqsdb.from_sequences([b"MKV*MAA", b"AAAA", b"*MKK"], 0, keys=[3, 7, 9])
    ddb = ctx.upload(seq)
    out, add_stop = ctx.findassemblystart(ddb, np.zeros(0, dtype=api.ALN))
    got = out.download()
    out.free(); ddb.free()
    assert (add_stop == -1).all()
    assert got.entries_by_key() == seq.entries_by_key()


def test_cyclecheck_matches_oracle_and_golden(golden_root, ctx):
    d, man = golden_case("cycle_nt", golden_root)
    seq = mmseqsdb.read_db(os.path.join(d, "seqs"))
    ddb = ctx.upload(seq)
    for s in man["steps"]:
        flags = parse_flags(s["args"])
        split = ctx.cyclecheck(ddb, int(flags["--max-seq-len"]))
        want = ob.cyclecheck(seq, int(flags["--max-seq-len"]))
        assert np.array_equal(split, want), np.nonzero(split != want)[0][:10]
        golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][1]))
        assert_same_entries(ob.cycle_db(seq, split, int(flags["--chop-cycle"])).entries_by_key(), golden.entries_by_key(), "cycle_nt/%s" % s["dbs"][1])
    ddb.free()


@pytest.mark.parametrize("case", ["synth_nt", "long_nt"])
def test_cyclecheck_on_assemblies_matches_oracle(case, golden_root, ctx):
    """The DBs cyclecheck sees inside the workflow (nuclassemble.sh:19-60), up to 80 000 nt (CTA-per-sequence path)."""
    d, man = golden_case(case, golden_root)
    for name in sorted(f for f in os.listdir(d) if f.endswith("_noneCycle")):
        seq = mmseqsdb.read_db(os.path.join(d, name[: -len("_noneCycle")]))
        none = mmseqsdb.read_db(os.path.join(d, name))
        ddb = ctx.upload(seq)
        split = ctx.cyclecheck(ddb, 200000)
        ddb.free()
        assert np.array_equal(split, ob.cyclecheck(seq, 200000)), (case, name)
        assert sorted(int(k) for k in seq.keys[split == 0]) == sorted(int(k) for k in none.keys), (case, name)


def test_cyclecheck_random_lengths_match_oracle(ctx):
    """Lengths around the boundaries of the three kernel classes (380 / 381, 1532 / 1533), below the k-mer size, with N
    runs and with every kind of self overlap."""
    rng = np.random.default_rng(33)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for L in [1, 5, 21, 22, 23, 65, 66, 67, 379, 380, 381, 382, 1531, 1532, 1533, 1534, 3000, 5000] + [int(x) for x in rng.integers(30, 2500, 300)]:
        g = acgt[rng.integers(0, 4, max(1, int(L * rng.uniform(0.4, 1.0))))]
        s = np.concatenate([g, g])[:L].copy() if rng.random() < 0.7 else acgt[rng.integers(0, 4, L)]
        if len(s) < L:
            s = np.concatenate([s, acgt[rng.integers(0, 4, L - len(s))]])
        if rng.random() < 0.2 and L > 40:
            p = int(rng.integers(0, L - 5)); s[p:p + 3] = ord("N")
        seqs.append(s.tobytes())
    seq = mmseqsdb.from_sequences(seqs, 1)
    ddb = ctx.upload(seq)
    for max_len in (200000, 1000):
        split = ctx.cyclecheck(ddb, max_len)
        want = ob.cyclecheck(seq, max_len)
        assert np.array_equal(split, want), (max_len, np.nonzero(split != want)[0][:10])
    ddb.free()
    assert (want > 0).sum() > 20


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, "%s\n%s" % (" ".join(cmd), r.stdout[-3000:])


def test_cli_findassemblystart_and_cyclecheck(golden_root, tmp_path):
    d, man = golden_case("synth_aa", golden_root)
    out = str(tmp_path / "corrected_gpu")
    run([CLI, "findassemblystart", os.path.join(d, "aa_6f_start_long"), os.path.join(d, "aln_0"), out, "--threads", "4", "-v", "3"])
    got, want = mmseqsdb.read_db(out), mmseqsdb.read_db(os.path.join(d, "corrected_seqs"))
    assert got.dbtype == want.dbtype
    assert_same_entries(got.entries_by_key(), want.entries_by_key(), "corrected_seqs via CLI")
    d, man = golden_case("cycle_nt", golden_root)
    for s in man["steps"]:
        out = str(tmp_path / (s["dbs"][1] + "_gpu"))
        run([CLI, "cyclecheck", os.path.join(d, "seqs"), out] + s["args"])
        got, want = mmseqsdb.read_db(out), mmseqsdb.read_db(os.path.join(d, s["dbs"][1]))
        assert got.dbtype == want.dbtype == 1
        assert_same_entries(got.entries_by_key(), want.entries_by_key(), "%s via CLI" % s["dbs"][1])


# ---- extractorfs / translatenucs / concatdbs ------------------------------------------------

def test_extractorfs_translatenucs_match_oracle_and_golden(golden_root, ctx):
    d, man = golden_case("orf_aa", golden_root)
    reads = mmseqsdb.read_db(os.path.join(d, "nucl_reads"))
    ddb = ctx.upload(reads)
    for s in [s for s in man["steps"] if s["cmd"] == "extractorfs"]:
        name = s["dbs"][1]
        flags = parse_flags(s["args"])
        op = ob.orf_params_from_flags(flags, api.OrfParams)
        wnuc, winfo = ob.extractorfs(reads, ob.orf_params_from_flags(flags), False)
        # nucleotide fragments + ORF header fields
        out, info = ctx.extractorfs(ddb, op, translate=False)
        got = out.download()
        assert np.array_equal(info, winfo), name
        assert_same_entries(got.entries_by_key(), wnuc.entries_by_key(), "orf_aa/%s vs oracle" % name)
        golden = mmseqsdb.read_db(os.path.join(d, name))
        assert_same_entries(got.entries_by_key(), golden.entries_by_key(), "orf_aa/%s vs reference" % name)
        assert_same_entries(ob.orf_header_entries(info), mmseqsdb.read_db(os.path.join(d, name + "_h")).entries_by_key(), "orf_aa/%s_h" % name)
        # translatenucs --add-orf-stop 1 on that DB (device resident), flags from the header fields
        tflags = ((info[:, 3] & 1) == 0).astype(np.uint8) | (((info[:, 3] & 2) == 0).astype(np.uint8) << 1)
        aa = ctx.translatenucs(out, tflags)
        gaa = aa.download()
        aa.free(); out.free()
        want_aa = mmseqsdb.read_db(os.path.join(d, "aa_" + name[len("nucl_"):]))
        assert_same_entries(gaa.entries_by_key(), want_aa.entries_by_key(), "orf_aa/translatenucs %s" % name)
        # the fused form
        out, info2 = ctx.extractorfs(ddb, op, translate=True)
        got = out.download()
        out.free()
        assert np.array_equal(info2, winfo)
        assert_same_entries(got.entries_by_key(), want_aa.entries_by_key(), "orf_aa/fused %s" % name)
    # translatenucs on the raw reads: all length classes mod 3, entries shorter than a codon are dropped
    aa = ctx.translatenucs(ddb)
    got = aa.download()
    aa.free(); ddb.free()
    assert_same_entries(got.entries_by_key(), mmseqsdb.read_db(os.path.join(d, "aa_reads")).entries_by_key(), "orf_aa/aa_reads")
    assert_same_entries(got.entries_by_key(), ob.translatenucs(reads).entries_by_key(), "orf_aa/aa_reads vs oracle")


def test_six_frame_fragments_equal_workflow_input(golden_root, ctx):
    """aa_6f_start_long = concatdbs(aa_6f_long, aa_6f_start) (data/assemble.sh:41-77) from the reads in one go."""
    d, man = golden_case("orf_aa", golden_root)
    reads = mmseqsdb.read_db(os.path.join(d, "nucl_reads"))
    ddb = ctx.upload(reads)
    out = ctx.six_frame_fragments(ddb)
    got = out.download()
    out.free(); ddb.free()
    lo, st = mmseqsdb.read_db(os.path.join(d, "aa_long")), mmseqsdb.read_db(os.path.join(d, "aa_start"))
    want = {int(k): lo.entry(i) for i, k in enumerate(lo.keys)}
    want.update({lo.n + int(k): st.entry(i) for i, k in enumerate(st.keys)})
    assert got.n == lo.n + st.n
    assert_same_entries(got.entries_by_key(), want, "aa_6f_start_long")


def test_extractorfs_empty_and_tiny(ctx):
    seq = mmseqsdb.from_sequences([b"A", b"AC", b"ATG", b"ATGAAATAG", b"NNNNNNNNNNNN"], 1, keys=[2, 5, 6, 9, 11])
    ddb = ctx.upload(seq)
    op = api.OrfParams(min_length=1, max_length=32734, max_gaps=2147483647, contig_start_mode=2, contig_end_mode=2, orf_start_mode=1,
                       forward_frames=7, reverse_frames=7, translation_table=1, use_all_table_starts=0)
    for tr in (False, True):
        out, info = ctx.extractorfs(ddb, op, translate=tr)
        got = out.download()
        out.free()
        want, winfo = ob.extractorfs(seq, ob.OrfParams(**{f: getattr(op, f) for f, _ in ob.OrfParams._fields_}), tr)
        assert np.array_equal(info, winfo)
        assert got.entries_by_key() == want.entries_by_key()
    ddb.free()


def test_cli_extractorfs_and_translatenucs(golden_root, tmp_path):
    d, man = golden_case("orf_aa", golden_root)
    for s in man["steps"]:
        if s["cmd"] == "extractorfs" and s["dbs"][1] in ("nucl_start", "nucl_long"):
            out = str(tmp_path / (s["dbs"][1] + "_gpu"))
            run([CLI, "extractorfs", os.path.join(d, "nucl_reads"), out] + s["args"])
            for ext in ("", "_h"):
                got, want = mmseqsdb.read_db(out + ext), mmseqsdb.read_db(os.path.join(d, s["dbs"][1] + ext))
                assert got.dbtype == want.dbtype
                assert_same_entries(got.entries_by_key(), want.entries_by_key(), "%s%s via CLI" % (s["dbs"][1], ext))
            aa = str(tmp_path / ("aa_" + s["dbs"][1] + "_gpu"))
            run([CLI, "translatenucs", out, aa, "--translation-table", "1", "--add-orf-stop", "1", "-v", "3", "--compressed", "0", "--threads", "1"])
            got, want = mmseqsdb.read_db(aa), mmseqsdb.read_db(os.path.join(d, "aa_" + s["dbs"][1][len("nucl_"):]))
            assert got.dbtype == want.dbtype == 0
            assert_same_entries(got.entries_by_key(), want.entries_by_key(), "aa_%s via CLI" % s["dbs"][1])
            assert os.path.exists(aa + "_h.index")          # the header DB travels with the translated DB (softlinkDb)
