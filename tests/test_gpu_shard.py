"""Multi-GPU data plane (-m gpu): pg_shard_iteration / pg_shard_allgather_db / pg_shard_broadcast_db over NCCL.

* world = 1 (one GPU, a one-rank communicator): the whole C++ plane -- count matrix, grouped send / recv, work histogram,
  equal-work bounds, all-gather -- must reproduce pg_assemble_iteration exactly;
* world = 2 (skipped unless two GPUs are visible): two torchrun ranks (tests/shard_worker.py) compare their shares with a
  single-GPU run of the same DB, chain two iterations through the all-gathered DB, aa and nt;
* hash-range splits on one GPU (--split-memory-limit): forced 2 / 4 / 7 splits must equal the unsplit run and the golden DB."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import golden_case, ROOT
from plass_b200 import mmseqsdb, api
import params
from test_oracle_vs_reference import hits_from_pref

pytestmark = pytest.mark.gpu


def case_inputs(case, golden_root):
    d, man = golden_case(case, golden_root)
    s = [x for x in man["steps"] if x["cmd"] == "kmermatcher"][0]
    rs = [x for x in man["steps"] if x["cmd"] == "rescorediagonal"][0]
    ex = [x for x in man["steps"] if x["cmd"] in ("assembleresults", "nuclassembleresults")][0]
    seq = mmseqsdb.read_db(os.path.join(d, s["dbs"][0]))
    nucl = seq.dbtype == 1
    kp = api.KmParams(hash_start=0, hash_end=65535, **params.km_fields(s["args"], nucl))
    return d, s, seq, kp, api.RsParams(**params.rs_fields(rs["args"])), api.ExParams(**params.ex_fields(ex["args"]))


def same_records(a, b, fields):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in fields)


HIT_F = ("rep", "target", "score", "diag")
ALN_F = ("query", "target", "bits", "seq_id", "evalue", "q_start", "q_end", "q_len", "db_start", "db_end", "db_len")


@pytest.mark.parametrize("case", ["synth_aa", "synth_nt", "example_aa"])
def test_shard_iteration_one_rank_equals_fused_iteration(case, golden_root):
    ctx = api.Context(0)
    try:
        ctx.comm_init(0, 1, api.Context.comm_unique_id())
        _, _, seq, kp, rp, ep = case_inputs(case, golden_root)
        ddb = ctx.upload(seq)
        ref_out, ref_hits, ref_alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
        ref_db = ref_out.download()
        rep = ctx.shard_broadcast_db(ddb, 0)                      # replica (trivial with one rank, same code path)
        out, own, hits, alns = ctx.shard_iteration(rep, kp, rp, ep, want_intermediates=True)
        assert own == (0, 0xFFFFFFFF)
        assert same_records(hits, ref_hits, HIT_F), case
        assert same_records(alns, ref_alns, ALN_F), case
        full = ctx.shard_allgather_db(out)
        got = full.download()
        for f in ("keys", "lens", "offsets", "data"):
            assert np.array_equal(getattr(got, f), getattr(ref_db, f)), (case, f)
        # chained second iteration from the all-gathered DB equals the single-GPU chain
        out2, _, hits2, _ = ctx.shard_iteration(full, kp, rp, ep, want_intermediates=True)
        ref2, ref_hits2, _ = ctx.assemble_iteration(ref_out, kp, rp, ep, want_intermediates=True)
        assert same_records(hits2, ref_hits2, HIT_F), case
        a, b = out2.download(), ref2.download()
        assert np.array_equal(a.data, b.data) and np.array_equal(a.lens, b.lens)
        for x in (out, out2, full, rep, ref_out, ref2, ddb):
            x.free()
    finally:
        ctx.close()


def test_shard_iteration_two_ranks(golden_root, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    for case in ("synth_aa", "synth_nt"):
        golden_case(case, golden_root)
    env = dict(os.environ, PLASS_GOLDEN_ROOT=str(golden_root), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "tests", "shard_worker.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=900)
    assert r.returncode == 0 and r.stdout.count("shard-ok") == 2, r.stdout[-4000:]


@pytest.mark.parametrize("case", ["synth_aa", "synth_nt", "example_aa"])
def test_hash_range_splits_equal_unsplit_run(case, golden_root):
    """--split-memory-limit (kmermatcher.cpp:608-624, :736-778): the kmermatcher stage in 2 / 4 / 7 hash-range splits and with a
    memory limit that forces splitting must give the unsplit hits (= the golden pref DB, up to the nt strand rule)."""
    ctx = api.Context(0)
    try:
        d, s, seq, kp, rp, ep = case_inputs(case, golden_root)
        nucl = seq.dbtype == 1
        ddb = ctx.upload(seq)
        ref = ctx.kmermatcher(ddb, kp)
        assert ctx.timings()["splits"] == 1
        if not nucl:
            want = hits_from_pref(mmseqsdb.read_db(os.path.join(d, s["dbs"][1])))
            assert same_records(ref, want, HIT_F)
        for n in (2, 4, 7):
            ctx.debug_force_splits(n)
            got = ctx.kmermatcher(ddb, kp)
            assert ctx.timings()["splits"] == n
            if nucl:
                # the first-group quirk (kmermatcher.cpp:463) applies to the smallest k-mer of every split, as in the reference's own split runs
                assert len(got) == len(ref) and all(np.array_equal(got[f], ref[f]) for f in ("rep", "target", "diag"))
                assert np.array_equal(np.abs(got["score"]), np.abs(ref["score"])) and int((got["score"] != ref["score"]).sum()) <= 64 * n
            else:
                assert same_records(got, ref, HIT_F), (case, n)
        ctx.debug_force_splits(0)
        # a limit of 1/3 of what the stage needs -> at least 4 splits, chosen by the library
        need = 2 * 16 * int(ctx.timings()["n_kmer_records"])
        ctx.set_split_memory_limit(max(need // 3, 1 << 22))
        out, hits, _ = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
        assert ctx.timings()["splits"] >= 2
        if not nucl:
            assert same_records(hits, ref, HIT_F)
        ctx.set_split_memory_limit(0)
        out.free(); ddb.free()
    finally:
        ctx.close()


def test_split_memory_limit_too_small_is_a_clean_error(golden_root):
    ctx = api.Context(0)
    try:
        _, _, seq, kp, _, _ = case_inputs("synth_aa", golden_root)
        ddb = ctx.upload(seq)
        ctx.set_split_memory_limit(1024)
        with pytest.raises(api.PlassGpuError):
            ctx.kmermatcher(ddb, kp)
        ctx.set_split_memory_limit(0)
        assert len(ctx.kmermatcher(ddb, kp)) > 0
        ddb.free()
    finally:
        ctx.close()
