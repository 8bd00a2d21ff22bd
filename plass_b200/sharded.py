"""Multi-GPU assemble iteration: one process per GPU, the sequence DB replicated in every HBM, the k-mer
hash space sharded over the ranks, ONE all-to-all of (rep, target, diagonal) pair records per iteration
(DESIGN.md §5).  torch.distributed (NCCL over NVLink) is only the transport: the records
that cross the links are produced and consumed by the CUDA kernels of libplassgpu.so."""
import ctypes as C
import time

import numpy as np

from . import api

REC_BYTES = 16


def hash_range(rank, world):
    """Contiguous slice of the 16-bit k-mer hash space owned by `rank` (inclusive bounds), the same key the
    reference splits on (kmermatcher.cpp:736-778)."""
    lo = (65536 * rank) // world
    hi = (65536 * (rank + 1)) // world - 1
    return lo, hi


def owner_range(max_key, rank, world):
    lo, hi = C.c_uint32(), C.c_uint32()
    lib = api.load_library()
    lib.pg_shard_owner_range(C.c_uint32(max_key), C.c_int(rank), C.c_int(world), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def split_bytes(counts):
    return [int(c) * REC_BYTES for c in counts]


def shard_km_params(kp, rank, world):
    p = api.KmParams()
    C.memmove(C.byref(p), C.byref(kp), C.sizeof(api.KmParams))
    p.hash_start, p.hash_end = hash_range(rank, world)
    return p


class ShardedIteration:
    def __init__(self, ctx, dist, rank, world):
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self._t = {}
        self.last_d2h_bytes = 0

    def step(self, ddb, kp, rp, ep, download=False):
        import torch
        ctx, dist = self.ctx, self.dist
        counts = ctx.shard_pairs(ddb, shard_km_params(kp, self.rank, self.world), self.world)
        t1 = ctx.timings()
        n_send = sum(counts)
        send = torch.empty(max(n_send, 1) * REC_BYTES, dtype=torch.uint8, device="cuda")
        ctx.shard_export(send.data_ptr(), n_send)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        send_counts = torch.tensor(counts, dtype=torch.int64, device="cuda")
        recv_counts = torch.empty(self.world, dtype=torch.int64, device="cuda")
        dist.all_to_all_single(recv_counts, send_counts)
        rc = [int(x) for x in recv_counts.tolist()]
        n_recv = sum(rc)
        recv = torch.empty(max(n_recv, 1) * REC_BYTES, dtype=torch.uint8, device="cuda")
        dist.all_to_all_single(recv[: n_recv * REC_BYTES], send[: n_send * REC_BYTES], split_bytes(rc), split_bytes(counts))
        ev1.record()
        torch.cuda.synchronize()
        own = owner_range(ddb.max_key, self.rank, self.world)
        out, hits, alns = ctx.shard_finish(ddb, recv.data_ptr(), n_recv, own, rp, ep, want_intermediates=download)
        t2 = ctx.timings()
        t2["exchange_ms"] = ev0.elapsed_time(ev1)
        t2["total_ms"] = t2["total_ms"] + t2["exchange_ms"]
        self._t = t2
        if download:
            host = out.download()
            self.last_d2h_bytes = int(hits.nbytes + alns.nbytes + host.data.nbytes + host.offsets.nbytes + host.lens.nbytes + host.keys.nbytes)
            del hits, alns, host
        return out

    def timings(self):
        return self._t
