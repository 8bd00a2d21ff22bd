"""GPU parity tests (-m gpu) on the long-sequence case: sequences of up to 80 000 nt, where the reference switches to
KmerPosition<int> (kmermatcher.cpp:797-802, "wide" records) and rescorediagonal has to unwrap 16-bit diagonals
(DistanceCalculator.h:94-113).  Fixture: tests/golden/long_nt (penguin nuclassemble, two iterations)."""
import os
import numpy as np
import pytest

from common import golden_case
from plass_b200 import mmseqsdb, api
import oracle_binding as ob
import params
from test_oracle_vs_reference import hits_from_pref, alns_from_db, assert_same_entries
from test_gpu_parity import gpu_km, check_alns

pytestmark = pytest.mark.gpu

CASE = "long_nt"


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def test_wide_kmermatcher_matches_oracle_and_golden(golden_root, ctx):
    """Wide (T = int) records: rank-based 16-byte layout, 32-bit diagonals.  Integer columns bit-exact; for nucleotides the
    strand sign of a hit is not well defined in the reference when the records of the winning diagonal disagree
    (hazard 6), so a handful of sign-only differences is tolerated and nothing else."""
    d, man = golden_case(CASE, golden_root)
    for s in [s for s in man["steps"] if s["cmd"] == "kmermatcher"]:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        assert int(seq.lens.max()) >= 32767
        ddb = ctx.upload(seq)
        got = ctx.kmermatcher(ddb, gpu_km(s["args"], True))
        ddb.free()
        want = ob.kmermatch(seq, params.oracle_km(s["args"], True))
        assert len(got) == len(want), (s["dbs"][1], len(got), len(want))
        for f in ("rep", "target", "diag"):
            assert np.array_equal(got[f], want[f]), (s["dbs"][1], f)
        assert np.array_equal(np.abs(got["score"]), np.abs(want["score"])), (s["dbs"][1], "score")
        flipped = np.sign(got["score"]) != np.sign(want["score"])
        flips = int(flipped.sum())
        # hazard 6 concerns the first k-mer group only: every sign-only difference sits under that one representative
        assert len(np.unique(got["rep"][flipped])) <= 1, (s["dbs"][1], "strand flips under several representatives", flips)
        if flips == 0:
            golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][1]))
            assert_same_entries(ob.format_hits_by_rep(seq.keys, got), golden.entries_by_key(), "%s/%s" % (CASE, s["dbs"][1]))


def test_wide_rescorediagonal_matches_oracle_and_golden(golden_root, ctx):
    d, man = golden_case(CASE, golden_root)
    for s in [s for s in man["steps"] if s["cmd"] == "rescorediagonal"]:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        hits = hits_from_pref(mmseqsdb.read_db(os.path.join(d, s["dbs"][2])))
        ddb = ctx.upload(seq)
        got = ctx.rescorediagonal(ddb, hits, api.RsParams(**params.rs_fields(s["args"])))
        ddb.free()
        want = ob.rescore(seq, hits, params.oracle_rs(s["args"]))
        check_alns(got, want, "%s/%s" % (CASE, s["dbs"][3]))
        golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][3])).entries_by_key()
        text = ob.format_alns_by_query(seq.keys, got.astype(ob.ALN))
        bad = [k for k in golden if text[k] != golden[k]]
        assert len(bad) <= max(1, len(golden) // 10000), (s["dbs"][3], len(bad), text[bad[0]], golden[bad[0]])


def test_wide_nuclassembleresults_matches_oracle_and_golden(golden_root, ctx):
    d, man = golden_case(CASE, golden_root)
    for s in [s for s in man["steps"] if s["cmd"] == "nuclassembleresults"]:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        alns = alns_from_db(mmseqsdb.read_db(os.path.join(d, s["dbs"][1])))
        ddb = ctx.upload(seq)
        out, ext = ctx.assembleresults(ddb, alns, api.ExParams(**params.ex_fields(s["args"])))
        got = out.download()
        out.free(); ddb.free()
        want, wext = ob.extend(seq, alns, params.oracle_ex(s["args"]))
        assert np.array_equal(ext, wext), s["dbs"][2]
        assert_same_entries(got.entries_by_key(), want.entries_by_key(), "%s/%s vs oracle" % (CASE, s["dbs"][2]))
        golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][2]))
        assert_same_entries(got.entries_by_key(), golden.entries_by_key(), "%s/%s vs reference" % (CASE, s["dbs"][2]))


def test_wide_fused_iteration_is_consistent(golden_root, ctx):
    """Fused iteration: the alignments and contigs must be what the oracle's rescorediagonal and extension produce from the
    GPU's own prefilter hits (exact), and equal the reference's DBs when no strand sign differs."""
    d, man = golden_case(CASE, golden_root)
    steps = man["steps"]
    for i, s in enumerate(steps):
        if s["cmd"] != "nuclassembleresults":
            continue
        km = [x for x in steps[:i] if x["cmd"] == "kmermatcher" and x["dbs"][0] == s["dbs"][0]][-1]
        rs = [x for x in steps[:i] if x["cmd"] == "rescorediagonal" and x["dbs"][3] == s["dbs"][1]][-1]
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        ddb = ctx.upload(seq)
        out, hits, alns = ctx.assemble_iteration(ddb, gpu_km(km["args"], True), api.RsParams(**params.rs_fields(rs["args"])),
                                                 api.ExParams(**params.ex_fields(s["args"])), want_intermediates=True)
        got = out.download()
        out.free(); ddb.free()
        walns = ob.rescore(seq, np.ascontiguousarray(hits, dtype=ob.HIT), params.oracle_rs(rs["args"]))
        check_alns(alns, walns, "%s/%s fused" % (CASE, rs["dbs"][3]))
        wout, _ = ob.extend(seq, walns, params.oracle_ex(s["args"]))
        assert_same_entries(got.entries_by_key(), wout.entries_by_key(), "%s/%s fused vs oracle" % (CASE, s["dbs"][2]))
        whits = ob.kmermatch(seq, params.oracle_km(km["args"], True))
        if len(whits) == len(hits) and np.array_equal(whits["score"], hits["score"]):
            golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][2]))
            assert_same_entries(got.entries_by_key(), golden.entries_by_key(), "%s/%s fused vs reference" % (CASE, s["dbs"][2]))
