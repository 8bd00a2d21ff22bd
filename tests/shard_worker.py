"""Two-rank worker of tests/test_gpu_shard.py::test_shard_iteration_two_ranks (launched with torchrun, one GPU per rank)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from plass_b200 import api, sharded  # noqa: E402
from test_gpu_shard import case_inputs, HIT_F, ALN_F  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    runner = sharded.ShardedIteration(ctx, dist, rank, world)
    root = os.environ["PLASS_GOLDEN_ROOT"]
    for case in ("synth_aa", "synth_nt"):
        _, _, seq, kp, rp, ep = case_inputs(case, root)
        nucl = seq.dbtype == 1
        # rank 0 uploads, everybody gets the replica over NVLink
        ddb = runner.build_and_broadcast(lambda: ctx.upload(seq))
        assert ddb.n == seq.n
        ref_out, ref_hits, ref_alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)       # every rank: the single-GPU answer
        ref_db = ref_out.download()
        out, own, hits, alns = ctx.shard_iteration(ddb, kp, rp, ep, want_intermediates=True)
        lo, hi = own
        mh = (ref_hits["rep"] >= lo) & (ref_hits["rep"] < hi)
        ma = (ref_alns["query"] >= lo) & (ref_alns["query"] < hi)
        if nucl:
            # strand flag of the first k-mer group: the job-wide smallest k-mer is all-reduced, so this is exact too
            pass
        assert len(hits) == int(mh.sum()) and all(np.array_equal(hits[f], ref_hits[f][mh]) for f in HIT_F), (case, rank, "hits")
        assert len(alns) == int(ma.sum()) and all(np.array_equal(alns[f], ref_alns[f][ma]) for f in ALN_F), (case, rank, "alns")
        full = ctx.shard_allgather_db(out)
        got = full.download()
        for f in ("keys", "lens", "offsets", "data"):
            assert np.array_equal(getattr(got, f), getattr(ref_db, f)), (case, rank, f)
        # second iteration from the gathered DB (data/assemble.sh:153) against the single-GPU chain
        out2, own2, hits2, _ = ctx.shard_iteration(full, kp, rp, ep, want_intermediates=True)
        ref2, ref_hits2, _ = ctx.assemble_iteration(ref_out, kp, rp, ep, want_intermediates=True)
        m2 = (ref_hits2["rep"] >= own2[0]) & (ref_hits2["rep"] < own2[1])
        assert len(hits2) == int(m2.sum()) and all(np.array_equal(hits2[f], ref_hits2[f][m2]) for f in HIT_F), (case, rank, "hits of iteration 2")
        # sliced upload: PCIe carries this rank's slice only
        pinned = runner.pinned_slice(ddb)
        again = runner.upload_sliced(pinned)
        g2 = again.download()
        h0 = ddb.download()
        for f in ("keys", "lens", "offsets", "data"):
            assert np.array_equal(getattr(g2, f), getattr(h0, f)), (case, rank, "sliced upload", f)
        chk = runner.verify_against_single_gpu(ddb, kp, rp, ep)
        if rank == 0:
            assert chk["equal"], chk
        for x in (out, out2, full, ref_out, ref2, again, ddb):
            x.free()
    print("shard-ok rank %d" % rank, flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
