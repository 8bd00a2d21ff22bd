"""world_size-2 gloo worker for tests/test_host_cpu.py::test_sharding_plan_two_ranks_gloo: exercises the host
side of the multi-GPU plan (hash ranges, owner ranges, all-to-all split sizes) without a GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plass_b200 import sharded  # noqa: E402


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # hash ranges tile [0, 65535] without gaps or overlap
    rs = [sharded.hash_range(r, world) for r in range(world)]
    assert rs[0][0] == 0 and rs[-1][1] == 65535 and all(rs[i][1] + 1 == rs[i + 1][0] for i in range(world - 1))
    # owner ranges tile the key space and agree with the device-side rule owner = min(world-1, key / per)
    max_key = 1000003
    own = [sharded.owner_range(max_key, r, world) for r in range(world)]
    per = (max_key + world) // world
    keys = np.random.default_rng(rank).integers(0, max_key + 1, 10000)
    for k in keys:
        o = min(world - 1, int(k) // per)
        assert own[o][0] <= k < own[o][1]
    assert own[0][0] == 0 and own[-1][1] == 0xFFFFFFFF
    # exchange of synthetic pair records: counts first, then variable-sized payload; check conservation
    rng = np.random.default_rng(100 + rank)
    counts = [int(x) for x in rng.integers(0, 50, world)]
    payload = torch.from_numpy(np.concatenate([np.full(c * sharded.REC_BYTES, 16 * rank + d, dtype=np.uint8) for d, c in enumerate(counts)] + [np.zeros(0, np.uint8)]))
    sc, rc = torch.tensor(counts, dtype=torch.int64), torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(rc, sc)
    rcl = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(rcl) * sharded.REC_BYTES, dtype=torch.uint8)
    dist.all_to_all_single(recv, payload, sharded.split_bytes(rcl), sharded.split_bytes(counts))
    off = 0
    for src, c in enumerate(rcl):
        blk = recv[off: off + c * sharded.REC_BYTES]
        assert bool((blk == 16 * src + rank).all())
        off += c * sharded.REC_BYTES
    # k-mer owner bins: 256 hash bins map monotonically onto the ranks, every rank gets at least one bin (world <= 256)
    own_bins = [b * world // 256 for b in range(256)]
    assert own_bins[0] == 0 and own_bins[-1] == world - 1 and all(0 <= own_bins[i + 1] - own_bins[i] <= 1 for i in range(255))
    # sliced upload: every rank contributes its slice of the DB, the broadcasts reassemble the whole DB everywhere
    from plass_b200 import synth
    db = synth.protein_fragments(synth.make_reads(301, seed=5))
    sb = sharded.slice_bounds(db.n, world)
    assert sb[0][0] == 0 and sb[-1][1] == db.n and all(sb[i][1] == sb[i + 1][0] for i in range(world - 1))
    (g_data, g_offs, g_lens, g_keys), h2d = sharded.gather_slices(dist, db, rank, world, torch.device("cpu"))
    assert np.array_equal(g_data[:-16].numpy(), np.asarray(db.data)) and np.array_equal(g_offs[:-1].numpy().view(np.uint64), np.asarray(db.offsets, dtype=np.uint64))
    assert np.array_equal(g_lens[:-1].numpy().view(np.uint32), np.asarray(db.lens, dtype=np.uint32)) and np.array_equal(g_keys[:-1].numpy().view(np.uint32), np.asarray(db.keys, dtype=np.uint32))
    tb = torch.tensor([h2d], dtype=torch.int64)
    dist.all_reduce(tb)
    assert int(tb[0]) == db.data.nbytes + 16 * db.n
    # balanced representative ranges: every rank derives the same bounds from the all-reduced histogram, and the
    # ranges carry (nearly) equal work although the histogram is heavily skewed towards low keys
    hist = (np.random.default_rng(7 + rank).random(4096) * 1000 * np.exp(-np.arange(4096) / 500.0)).astype(np.int64)
    th = torch.from_numpy(hist.copy())
    dist.all_reduce(th)
    bounds = sharded.balanced_bounds(th.numpy(), max_key, world, per_key_weight=0.0)
    gathered = [None] * world
    dist.all_gather_object(gathered, bounds)
    assert all(g == bounds for g in gathered)
    assert bounds[0] == 0 and bounds[-1] == 0xFFFFFFFF and all(bounds[i] <= bounds[i + 1] for i in range(world))
    span = max_key + 1
    bin_of = lambda key: min(4095, key * 4096 // span)
    tot_h = th.numpy().astype(np.float64)
    share = [tot_h[bin_of(bounds[r]): (bin_of(bounds[r + 1]) if r + 1 < world else 4096)].sum() / tot_h.sum() for r in range(world)]
    assert all(abs(x - 1.0 / world) < 0.02 for x in share), share
    tot = torch.tensor([sum(counts), sum(rcl)], dtype=torch.int64)
    dist.all_reduce(tot)
    assert int(tot[0]) == int(tot[1])
    print("gloo-ok rank %d" % rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
