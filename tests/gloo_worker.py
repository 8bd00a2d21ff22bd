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
    tot = torch.tensor([sum(counts), sum(rcl)], dtype=torch.int64)
    dist.all_reduce(tot)
    assert int(tot[0]) == int(tot[1])
    print("gloo-ok rank %d" % rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
