"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (libplassgpu.so), against the CPU
oracle on the same inputs and against the committed golden fixtures of the reference binary.
Bar: bit-exact for every integer field; seq.id exact as float; E-value within 1e-6 relative
(BASELINE.json north_star), and the printed 10-column alignment text must match where the E-value
prints identically."""
import os
import numpy as np
import pytest

from common import golden_case
from plass_b200 import mmseqsdb, api
import oracle_binding as ob
import params
from test_oracle_vs_reference import hits_from_pref, alns_from_db, assert_same_entries

pytestmark = pytest.mark.gpu

CASES = ["example_aa", "synth_aa", "synth_nt"]


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def gpu_km(args, nucl):
    return api.KmParams(hash_start=0, hash_end=65535, **params.km_fields(args, nucl))


def test_radix_sort_matches_numpy(ctx):
    rng = np.random.default_rng(5)
    for n in (1, 31, 4096, 4097, 100003, 1 << 20):
        recs = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.uint64)
        recs[:, 0] &= np.uint64((1 << 40) - 1)
        recs[: n // 2, 0] &= np.uint64(0xFF)          # heavy duplicates exercise stability
        got = ctx.debug_radix_sort(recs.copy(), [(0, 0, 40)])
        order = np.argsort(recs[:, 0], kind="stable")
        assert np.array_equal(got, recs[order]), n
    # two-word key: (w0 bits 0..20 major, w1 bits 0..16 minor)
    recs = rng.integers(0, 1 << 63, size=(300000, 2), dtype=np.uint64)
    got = ctx.debug_radix_sort(recs.copy(), [(1, 0, 16), (0, 0, 20)])
    key = ((recs[:, 0] & np.uint64((1 << 20) - 1)) << np.uint64(16)) | (recs[:, 1] & np.uint64(0xFFFF))
    assert np.array_equal(got, recs[np.argsort(key, kind="stable")])


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_radix_pass_variants(mode, golden_root, ctx):
    """The 256-bin pass kernels -- register tile (round 1) and the persistent bulk-copy (TMA) variants (512 threads x 6 records /
    2 stages, 256 x 8 / 2 stages at 3 CTAs per SM, 512 x 4 / 3 stages) -- sort identically (stable), at sizes around the tile edges, and the
    kmermatcher (whose last partition pass also emits the bucket bounds in the bulk-copy variants) reproduces the golden hits."""
    lib = api.load_library()
    before = lib.pg_debug_get_radix_mode()
    lib.pg_debug_set_radix_mode(mode)
    try:
        rng = np.random.default_rng(50 + mode)
        for n in (1, 17, 2047, 2048, 2049, 3071, 3072, 3073, 6144, 100003, (1 << 21) + 5):
            recs = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.uint64)
            recs[: n // 2, 0] &= np.uint64(0xFF)          # heavy duplicates exercise stability
            got = ctx.debug_radix_sort(recs.copy(), [(0, 0, 24)])
            key = recs[:, 0] & np.uint64((1 << 24) - 1)
            assert np.array_equal(got, recs[np.argsort(key, kind="stable")]), (mode, n)
        recs = rng.integers(0, 1 << 63, size=(700001, 2), dtype=np.uint64)
        got = ctx.debug_radix_sort(recs.copy(), [(1, 0, 16), (0, 0, 20)])
        key = ((recs[:, 0] & np.uint64((1 << 20) - 1)) << np.uint64(16)) | (recs[:, 1] & np.uint64(0xFFFF))
        assert np.array_equal(got, recs[np.argsort(key, kind="stable")])
        for case in CASES:
            _kmermatcher_case(case, golden_root, ctx)
    finally:
        lib.pg_debug_set_radix_mode(before)


def test_partition_histograms_counted_by_the_extraction(ctx, monkeypatch):
    """Sort #1's digit histograms are a by-product of the extraction kernels (no histogram sweep over the records).  A context
    that runs the sweep instead (PLASS_B200_NO_PREHIST=1) must give the same hits, aa and nt, at record counts that need one,
    two and three partition passes (the last pass's narrower digit is a fold of its 256 counted bins)."""
    from plass_b200 import synth
    cases = []
    for n_reads, seed in ((1500, 11), (40000, 12), (700000, 13)):      # 1, 2 and 3 partition passes
        cases.append((synth.protein_fragments(synth.make_reads(n_reads, seed=seed)), api.default_km_params(False)))
    cases.append((synth.nucleotide_db(synth.make_reads(60000, seed=14)), api.default_km_params(True)))
    got = []
    for db, kp in cases:
        ddb = ctx.upload(db)
        got.append(ctx.kmermatcher(ddb, kp).copy())
        ddb.free()
    monkeypatch.setenv("PLASS_B200_NO_PREHIST", "1")
    other = api.Context(0)
    try:
        for (db, kp), g in zip(cases, got):
            ddb = other.upload(db)
            want = other.kmermatcher(ddb, kp)
            ddb.free()
            assert len(g) == len(want) and len(want) > 0
            for f in ("rep", "target", "score", "diag"):
                assert np.array_equal(g[f], want[f]), f
    finally:
        other.close()


def test_spill_list_of_oversized_buckets_matches_oracle(ctx):
    """A k-mer that occurs in thousands of sequences (here: 1500x coverage of a 300 nt region inside a 20x data set) makes
    its bucket larger than the shared-memory hash join takes.  Only those buckets go through the spill list (copied out,
    sorted in full, grouped by the tile kernel); the iteration keeps the fast path and equals the oracle."""
    from plass_b200 import synth
    base = synth.make_reads(30000, coverage=20.0, seed=5)
    rng = np.random.default_rng(6)
    region = synth.make_genome(330, rng)
    starts = rng.integers(0, len(region) - 150 + 1, 3000)
    hot = region[starts[:, None] + np.arange(150)[None, :]]
    reads = np.concatenate([base, hot])
    db = synth.protein_fragments(reads)
    kp, rp, ep = api.default_km_params(False), api.default_rs_params(False), api.default_ex_params(False)
    ddb = ctx.upload(db)
    out, hits, alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
    t = ctx.timings()
    got = out.download()
    out.free(); ddb.free()
    assert 0 < t["spilled_records"] < t["n_kmer_records"] // 2, t
    okp = ob.KmParams(**{f: getattr(kp, f) for f, _ in ob.KmParams._fields_ if f not in ("hash_start", "hash_end")})
    okp.hash_start, okp.hash_end = 0, ob.U64MAX
    whits = ob.kmermatch(db, okp)
    assert len(hits) == len(whits) and all(np.array_equal(hits[f], whits[f]) for f in ("rep", "target", "score", "diag"))
    walns = ob.rescore(db, whits, ob.RsParams(**{f: getattr(rp, f) for f, _ in ob.RsParams._fields_}))
    check_alns(alns, walns, "spill list")
    wout, _ = ob.extend(db, walns, ob.ExParams(**{f: getattr(ep, f) for f, _ in ob.ExParams._fields_}))
    assert_same_entries(got.entries_by_key(), wout.entries_by_key(), "spill list")
    print("spill list: %d of %d k-mer records, %d hits" % (t["spilled_records"], t["n_kmer_records"], len(hits)))


@pytest.mark.parametrize("bits", [9, 10])
def test_wide_digit_radix_and_kmermatcher(bits, golden_root, ctx):
    """512- / 1024-bin radix passes (radix_scatter_wide_kernel): sort results and the whole kmermatcher stay identical."""
    lib = api.load_library()
    assert lib.pg_debug_set_digit_bits(ctx.handle, bits) == 0
    try:
        rng = np.random.default_rng(7)
        for n in (1, 33, 3072, 3073, 100003, 1 << 20):
            recs = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.uint64)
            recs[: n // 2, 0] &= np.uint64(0x3FF)          # heavy duplicates exercise stability
            got = ctx.debug_radix_sort(recs.copy(), [(0, 0, 40)])
            key = recs[:, 0] & np.uint64((1 << 40) - 1)
            assert np.array_equal(got, recs[np.argsort(key, kind="stable")]), (bits, n)
        recs = rng.integers(0, 1 << 63, size=(300000, 2), dtype=np.uint64)
        got = ctx.debug_radix_sort(recs.copy(), [(1, 0, 16), (0, 32, 55)])
        key = (((recs[:, 0] >> np.uint64(32)) & np.uint64((1 << 23) - 1)) << np.uint64(16)) | (recs[:, 1] & np.uint64(0xFFFF))
        assert np.array_equal(got, recs[np.argsort(key, kind="stable")])
        for case in CASES:
            _kmermatcher_case(case, golden_root, ctx)
    finally:
        lib.pg_debug_set_digit_bits(ctx.handle, 8)


@pytest.mark.parametrize("case", CASES)
def test_extract_matches_oracle(case, golden_root, ctx):
    d, man = golden_case(case, golden_root)
    for s in [s for s in man["steps"] if s["cmd"] == "kmermatcher"]:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        nucl = seq.dbtype == 1
        want = ob.extract_kmers(seq, params.oracle_km(s["args"], nucl))
        ddb = ctx.upload(seq)
        got = ctx.debug_extract(ddb, gpu_km(s["args"], nucl))
        ddb.free()
        w = np.zeros(len(want), dtype=got.dtype)
        w["w0"] = want["kmer"]
        w["w1"] = (want["id"].astype(np.uint64) << np.uint64(32)) | ((want["seq_len"].astype(np.uint64) & np.uint64(0xFFFF)) << np.uint64(16)) | (want["pos"].astype(np.uint64) & np.uint64(0xFFFF))
        assert len(got) == len(w), (case, s["dbs"][0])
        assert np.array_equal(np.sort(got, order=["w0", "w1"]), np.sort(w, order=["w0", "w1"])), (case, s["dbs"][0])


@pytest.mark.parametrize("full_sort", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_kmermatcher_matches_oracle_and_golden(case, full_sort, golden_root, ctx):
    """full_sort = 0: partial-key partition + shared-memory hash join (default); 1: 8-pass sort + group_kernel (fallback)."""
    api.load_library().pg_debug_force_full_sort(ctx.handle, full_sort)
    try:
        _kmermatcher_case(case, golden_root, ctx)
    finally:
        api.load_library().pg_debug_force_full_sort(ctx.handle, 0)


def _kmermatcher_case(case, golden_root, ctx):
    d, man = golden_case(case, golden_root)
    for s in [s for s in man["steps"] if s["cmd"] == "kmermatcher"]:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        nucl = seq.dbtype == 1
        ddb = ctx.upload(seq)
        got = ctx.kmermatcher(ddb, gpu_km(s["args"], nucl))
        ddb.free()
        want = ob.kmermatch(seq, params.oracle_km(s["args"], nucl))
        assert len(got) == len(want), (case, s["dbs"][1], len(got), len(want))
        for f in ("rep", "target", "score", "diag"):
            assert np.array_equal(got[f], want[f]), (case, s["dbs"][1], f)
        golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][1]))
        assert_same_entries(ob.format_hits_by_rep(seq.keys, got), golden.entries_by_key(), "%s/%s" % (case, s["dbs"][1]))


def check_alns(got, want, what):
    assert len(got) == len(want), (what, len(got), len(want))
    for f in ("query", "target", "bits", "q_start", "q_end", "q_len", "db_start", "db_end", "db_len"):
        assert np.array_equal(got[f], want[f]), (what, f)
    assert np.array_equal(got["seq_id"], want["seq_id"]), (what, "seq_id")          # float32, exact
    rel = np.abs(got["evalue"] - want["evalue"]) / np.maximum(np.abs(want["evalue"]), 1e-300)
    assert rel.max() <= 1e-6, (what, "evalue", rel.max())                             # north_star tolerance


@pytest.mark.parametrize("case", CASES)
def test_rescorediagonal_matches_oracle_and_golden(case, golden_root, ctx):
    d, man = golden_case(case, golden_root)
    for s in [s for s in man["steps"] if s["cmd"] == "rescorediagonal"]:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        hits = hits_from_pref(mmseqsdb.read_db(os.path.join(d, s["dbs"][2])))
        ddb = ctx.upload(seq)
        got = ctx.rescorediagonal(ddb, hits, api.RsParams(**params.rs_fields(s["args"])))
        ddb.free()
        want = ob.rescore(seq, hits, params.oracle_rs(s["args"]))
        check_alns(got, want, "%s/%s" % (case, s["dbs"][3]))
        golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][3])).entries_by_key()
        text = ob.format_alns_by_query(seq.keys, got.astype(ob.ALN))
        bad = [k for k in golden if text[k] != golden[k]]
        # the text can only differ where a last-ulp E-value difference flips the 4th significant digit
        assert len(bad) <= max(1, len(golden) // 10000), (case, s["dbs"][3], len(bad), text[bad[0]], golden[bad[0]])


@pytest.mark.parametrize("case", CASES)
def test_assembleresults_matches_oracle_and_golden(case, golden_root, ctx):
    d, man = golden_case(case, golden_root)
    for s in [s for s in man["steps"] if s["cmd"] in ("assembleresults", "nuclassembleresults")]:
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        alns = alns_from_db(mmseqsdb.read_db(os.path.join(d, s["dbs"][1])))
        ddb = ctx.upload(seq)
        out, ext = ctx.assembleresults(ddb, alns, api.ExParams(**params.ex_fields(s["args"])))
        got = out.download()
        out.free(); ddb.free()
        want, wext = ob.extend(seq, alns, params.oracle_ex(s["args"]))
        assert np.array_equal(ext, wext), (case, s["dbs"][2])
        assert_same_entries(got.entries_by_key(), want.entries_by_key(), "%s/%s vs oracle" % (case, s["dbs"][2]))
        golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][2]))
        assert_same_entries(got.entries_by_key(), golden.entries_by_key(), "%s/%s vs reference" % (case, s["dbs"][2]))


@pytest.mark.parametrize("case", CASES)
def test_fused_iteration_matches_golden_chain(case, golden_root, ctx):
    """kmermatcher -> rescorediagonal -> assembleresults without leaving HBM == the reference's three DBs."""
    d, man = golden_case(case, golden_root)
    steps = man["steps"]
    for i, s in enumerate(steps):
        if s["cmd"] not in ("assembleresults", "nuclassembleresults"):
            continue
        km = [x for x in steps[:i] if x["cmd"] == "kmermatcher" and x["dbs"][0] == s["dbs"][0]][-1]
        rs = [x for x in steps[:i] if x["cmd"] == "rescorediagonal" and x["dbs"][3] == s["dbs"][1]][-1]
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        nucl = seq.dbtype == 1
        ddb = ctx.upload(seq)
        out, hits, alns = ctx.assemble_iteration(ddb, gpu_km(km["args"], nucl), api.RsParams(**params.rs_fields(rs["args"])),
                                                 api.ExParams(**params.ex_fields(s["args"])), want_intermediates=True)
        got = out.download()
        out.free(); ddb.free()
        pref = mmseqsdb.read_db(os.path.join(d, km["dbs"][1]))
        assert_same_entries(ob.format_hits_by_rep(seq.keys, hits), pref.entries_by_key(), "%s/%s fused" % (case, km["dbs"][1]))
        golden = mmseqsdb.read_db(os.path.join(d, s["dbs"][2]))
        assert_same_entries(got.entries_by_key(), golden.entries_by_key(), "%s/%s fused" % (case, s["dbs"][2]))
        t = ctx.timings()
        assert t["kernel_launches"] > 0 and t["n_hits"] == len(hits)


def ctx_n_records(ctx, ddb, kp):
    recs = ctx.debug_extract(ddb, kp)
    return len(recs)


def ptr_n(recv_n):
    recv, n = recv_n
    ptr_n.keep = recv          # keep the tensor alive across the call that consumes its pointer
    return recv.data_ptr(), n


def emulated_all_to_all(sends, counts, dst):
    import torch
    parts = []
    for src in range(len(sends)):
        off = sum(counts[src][:dst]) * 16
        parts.append(sends[src][off: off + counts[src][dst] * 16])
    recv = torch.cat(parts)
    n = recv.numel() // 16
    if n == 0:
        recv = torch.empty(16, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    return recv, n


@pytest.mark.parametrize("case,world,mode", [("synth_aa", 2, "two"), ("synth_aa", 4, "two"), ("synth_nt", 2, "two"), ("synth_nt", 3, "two"),
                                             ("synth_aa", 2, "hash"), ("synth_nt", 2, "hash")])
def test_sharded_iteration_equals_single_gpu(case, world, mode, golden_root, ctx):
    """The multi-GPU decompositions, emulated rank by rank on one GPU: the union of the ranks' results must equal
    the unsharded run (and therefore the reference).  "two": sequence-sliced extraction -> all-to-all of k-mer
    records -> group -> all-to-all of pair records -> owner-local sort #2, rescoring, extension;  "hash": the
    reference's hash-range split of the extraction + only the pair exchange."""
    import torch
    from plass_b200 import sharded
    d, man = golden_case(case, golden_root)
    s = [x for x in man["steps"] if x["cmd"] == "kmermatcher"][0]
    rs = [x for x in man["steps"] if x["cmd"] == "rescorediagonal"][0]
    ex = [x for x in man["steps"] if x["cmd"] in ("assembleresults", "nuclassembleresults")][0]
    seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
    nucl = seq.dbtype == 1
    kp, rp, ep = gpu_km(s["args"], nucl), api.RsParams(**params.rs_fields(rs["args"])), api.ExParams(**params.ex_fields(ex["args"]))
    ddb = ctx.upload(seq)
    ref_out, ref_hits, ref_alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
    ref_db = ref_out.download()
    ref_out.free()
    def export(c):
        t = torch.empty(max(sum(c), 1) * 16, dtype=torch.uint8, device="cuda")
        ctx.shard_export(t.data_ptr(), sum(c))
        return t

    sends, counts = [], []
    if mode == "hash":
        for r in range(world):
            c = ctx.shard_pairs(ddb, sharded.shard_km_params(kp, r, world), world)
            sends.append(export(c)); counts.append(c)
    else:
        ksends, kcounts = [], []
        for r in range(world):
            c = ctx.shard_extract(ddb, kp, r, world)
            ksends.append(export(c)); kcounts.append(c)
        assert sum(sum(c) for c in kcounts) == ctx_n_records(ctx, ddb, kp)
        # the ranks run one after the other on one context, so the group phase runs twice: once for the summed
        # histogram (the all-reduce), once more right before each rank's route
        hist = sum(ctx.shard_group(ddb, kp, *ptr_n(emulated_all_to_all(ksends, kcounts, r))).astype(np.float64) for r in range(world))
        bounds = sharded.balanced_bounds(hist, ddb.max_key, world)
        assert bounds[0] == 0 and bounds[-1] == 0xFFFFFFFF and all(bounds[i] <= bounds[i + 1] for i in range(world))
        for r in range(world):
            h = ctx.shard_group(ddb, kp, *ptr_n(emulated_all_to_all(ksends, kcounts, r)))
            c = ctx.shard_route(bounds)
            assert sum(c) == int(h.sum())
            sends.append(export(c)); counts.append(c)
    if mode == "hash":
        bounds = [sharded.owner_range(ddb.max_key, r, world)[0] for r in range(world)] + [0xFFFFFFFF]
    all_hits, all_alns, entries = [], [], {}
    for dst in range(world):
        recv, n = emulated_all_to_all(sends, counts, dst)
        out, hits, alns = ctx.shard_finish(ddb, recv.data_ptr(), n, (bounds[dst], bounds[dst + 1]), rp, ep, want_intermediates=True)
        got = out.download()
        out.free()
        all_hits.append(hits.copy()); all_alns.append(alns.copy())
        e = got.entries_by_key()
        assert not (set(e) & set(entries))
        entries.update(e)
    ddb.free()
    hits, alns = np.concatenate(all_hits), np.concatenate(all_alns)
    assert len(hits) == len(ref_hits) and all(np.array_equal(hits[f], ref_hits[f]) for f in ("rep", "target", "score", "diag"))
    check_alns(alns, ref_alns, "%s sharded x%d" % (case, world))
    assert_same_entries(entries, ref_db.entries_by_key(), "%s sharded x%d output DB" % (case, world))


SWEEP = [
    # (case, overrides)  -- parameter values outside the workflow defaults exercise the generic kernel instances
    ("example_aa", dict(kmer_size=10, alph_size=21, kmers_per_seq=20)),
    ("example_aa", dict(kmer_size=12, kmers_per_seq=8, hash_shift=70, include_only_extendable=1)),
    ("example_aa", dict(ignore_multi_kmer=0, kmers_per_seq=25)),
    ("example_aa", dict(cov_mode=1, cov_thr=0.8)),
    ("synth_nt", dict(kmer_size=15, kmers_per_seq=30, kmers_per_seq_scale=0.2, include_only_extendable=0)),
    ("synth_nt", dict(kmer_size=27, kmers_per_seq=10, kmers_per_seq_scale=0.0, hash_shift=5)),
    ("synth_nt", dict(ignore_multi_kmer=0)),
]


@pytest.mark.parametrize("case,over", SWEEP)
def test_kmermatcher_param_sweep_matches_oracle(case, over, golden_root, ctx):
    d, man = golden_case(case, golden_root)
    steps = [s for s in man["steps"] if s["cmd"] == "kmermatcher"]
    for s in (steps[0], steps[-1]):           # first iteration (reads) and last (contigs of mixed length)
        seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
        nucl = seq.dbtype == 1
        f = params.km_fields(s["args"], nucl)
        f.update(over)
        okp = ob.KmParams(**f)
        okp.hash_start, okp.hash_end = 0, ob.U64MAX
        want = ob.kmermatch(seq, okp)
        ddb = ctx.upload(seq)
        got = ctx.kmermatcher(ddb, api.KmParams(hash_start=0, hash_end=65535, **f))
        ddb.free()
        assert len(got) == len(want), (case, over, len(got), len(want))
        for fld in ("rep", "target", "diag"):
            assert np.array_equal(got[fld], want[fld]), (case, over, fld)
        # the strand sign is only defined up to the documented tie hazard (nt); magnitudes must agree
        assert np.array_equal(np.abs(got["score"]), np.abs(want["score"])), (case, over)
        assert len(np.unique(got["rep"][got["score"] != want["score"]])) <= 1, (case, over)   # first k-mer group only


@pytest.mark.parametrize("coverage,n_reads", [(20, 4000), (60, 4000), (120, 4000), (500, 3000), (2500, 2500)])
def test_whole_iteration_at_high_coverage_matches_oracle(coverage, n_reads, ctx):
    """Representatives with hundreds to thousands of pair records and queries with more than 32 / 64 alignments:
    the (target, diagonal) aggregating reduce kernels (warp, CTA, full-sort spill), the wide extension paths and
    the fall-backs behind them, against the CPU oracle on the same seeded input."""
    from plass_b200 import synth
    reads = synth.make_reads(n_reads, coverage=float(coverage), seed=11 + coverage)
    db = synth.protein_fragments(reads)
    kp, rp, ep = api.default_km_params(False), api.default_rs_params(False), api.default_ex_params(False)
    ddb = ctx.upload(db)
    out, hits, alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
    got = out.download()
    out.free(); ddb.free()
    okp = ob.KmParams(**{f: getattr(kp, f) for f, _ in ob.KmParams._fields_ if f not in ("hash_start", "hash_end")})
    okp.hash_start, okp.hash_end = 0, ob.U64MAX
    whits = ob.kmermatch(db, okp)
    assert len(hits) == len(whits), (coverage, len(hits), len(whits))
    for f in ("rep", "target", "score", "diag"):
        assert np.array_equal(hits[f], whits[f]), (coverage, f)
    walns = ob.rescore(db, whits, ob.RsParams(**{f: getattr(rp, f) for f, _ in ob.RsParams._fields_}))
    check_alns(alns, walns, "coverage %d" % coverage)
    wout, _ = ob.extend(db, walns, ob.ExParams(**{f: getattr(ep, f) for f, _ in ob.ExParams._fields_}))
    assert_same_entries(got.entries_by_key(), wout.entries_by_key(), "coverage %d" % coverage)
    per_rep = np.bincount(whits["rep"]) if len(whits) else np.zeros(1)
    per_query = np.bincount(walns["query"]) if len(walns) else np.zeros(1)
    print("coverage %d: %d fragments, %d hits (max %d per representative), max %d alignments per query" % (
        coverage, db.n, len(whits), per_rep.max(), per_query.max()))


def test_async_upload_equals_blocking(golden_root, ctx):
    """pg_seqdb_upload_async: the copies are only enqueued (upload stream) and the DB is completed at its first use."""
    import torch
    d, man = golden_case("synth_aa", golden_root)
    s = [x for x in man["steps"] if x["cmd"] == "kmermatcher"][0]
    seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
    pinned = mmseqsdb.DB(torch.from_numpy(seq.data.copy()).pin_memory().numpy(), torch.from_numpy(seq.keys.copy()).pin_memory().numpy(),
                         torch.from_numpy(seq.offsets.view(np.int64).copy()).pin_memory().numpy().view(np.uint64),
                         torch.from_numpy(seq.lens.view(np.int32).copy()).pin_memory().numpy().view(np.uint32), seq.dbtype)
    kp = gpu_km(s["args"], False)
    a = ctx.upload(seq)
    want = ctx.kmermatcher(a, kp)
    a.free()
    b1, b2, b3 = ctx.upload_async(pinned), ctx.upload_async(pinned), ctx.upload_async(pinned)
    got1 = ctx.kmermatcher(b1, kp)
    back = b2.download()                       # first use = a download
    got2 = ctx.kmermatcher(b2, kp)
    b1.free(); b2.free(); b3.free()            # b3 is released without ever being used
    for f in ("rep", "target", "score", "diag"):
        assert np.array_equal(got1[f], want[f]) and np.array_equal(got2[f], want[f]), f
    assert back.entries_by_key() == seq.entries_by_key()


def test_async_results_equal_blocking(golden_root, ctx):
    """pg_set_async_results: two iterations in flight, results awaited by ticket, must equal the blocking calls."""
    from plass_b200 import synth
    dbs = [synth.protein_fragments(synth.make_reads(3000, seed=s)) for s in (3, 4, 5)]
    kp, rp, ep = api.default_km_params(False), api.default_rs_params(False), api.default_ex_params(False)
    want = []
    for db in dbs:
        ddb = ctx.upload(db)
        out, hits, alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
        want.append((hits.copy(), alns.copy(), out.download().entries_by_key()))
        out.free(); ddb.free()
    ctx.set_async_results(True)
    try:
        inflight = []
        for db in dbs:
            ddb = ctx.upload(db)
            out, hits, alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
            host = out.download()
            inflight.append((ctx.results_ticket(), hits, alns, host))
            out.free(); ddb.free()          # released while the copies may still be in flight
        for (ticket, hits, alns, host), (whits, walns, wout) in zip(inflight, want):
            ctx.results_wait(ticket)
            assert np.array_equal(hits, whits) and np.array_equal(alns, walns)
            assert host.entries_by_key() == wout
    finally:
        ctx.set_async_results(False)
