"""Multi-GPU assemble iteration: one process per GPU, the sequence DB replicated in every HBM.  Each rank
extracts the k-mers of its slice of the sequences; all-to-all #1 routes the k-mer records to the rank that owns
the k-mer (hash of the k-mer), which sorts and groups them; all-to-all #2 routes the (rep, target, diagonal) pair
records to the rank that owns the representative, which finishes kmermatcher, rescorediagonal and the extension
for its queries.  torch.distributed (NCCL over NVLink) is only the transport: the
records that cross the links are produced and consumed by the CUDA kernels of libplassgpu.so."""
import ctypes as C
import time

import numpy as np

from . import api

REC_BYTES = 16
_CONTIGUOUS_OK = set()


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


def balanced_bounds(hist, max_key, world, per_key_weight=4.0):
    """Cuts the representative key space [0, max_key] into `world` contiguous ranges of (nearly) equal work.
    hist[b] = pair records (summed over all ranks) whose representative falls into bin b = rep * BINS / (max_key+1);
    the work of a bin = its pair records + per_key_weight x its keys (every owned sequence is also a query with a self
    alignment and an output entry).  Returns world+1 ascending key bounds, bounds[0] = 0, bounds[-1] = 0xFFFFFFFF;
    cuts fall on bin edges, so every rank computes the same bounds from the same summed histogram."""
    hist = np.asarray(hist, dtype=np.float64)
    bins = len(hist)
    span = int(max_key) + 1
    edges = [(b * span + bins - 1) // bins for b in range(bins + 1)]       # smallest key of bin b (ceil), edges[bins] = span
    work = hist + per_key_weight * np.diff(np.asarray(edges, dtype=np.float64))
    cum = np.concatenate([[0.0], np.cumsum(work)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        b = int(np.searchsorted(cum, target, side="left"))                  # first edge with cum >= target
        b = min(max(b, 0), bins)
        bounds.append(max(edges[b], bounds[-1]))
    bounds.append(0xFFFFFFFF)
    return bounds


def split_bytes(counts):
    return [int(c) * REC_BYTES for c in counts]


def shard_km_params(kp, rank, world):
    p = api.KmParams()
    C.memmove(C.byref(p), C.byref(kp), C.sizeof(api.KmParams))
    p.hash_start, p.hash_end = hash_range(rank, world)
    return p


def slice_bounds(n, world):
    """Sequence index ranges [lo, hi) of the ranks: the slices pg_shard_extract works on."""
    return [((n * r) // world, (n * (r + 1)) // world) for r in range(world)]


def gather_slices(dist, db, rank, world, device):
    """Every rank copies only ITS slice of the four DB arrays to `device`; the slices are then all-gathered (one
    broadcast per source rank, the slices have different byte sizes).  Returns (data, offsets, lens, keys) tensors
    holding the whole DB and the bytes this rank copied from the host."""
    import torch
    n = int(db.n)
    offs = np.ascontiguousarray(db.offsets, dtype=np.uint64)
    lens = np.ascontiguousarray(db.lens, dtype=np.uint32)
    keys = np.ascontiguousarray(db.keys, dtype=np.uint32)
    data = np.ascontiguousarray(db.data)
    if id(db) not in _CONTIGUOUS_OK:      # O(n) host check, once per DB object
        assert n == 0 or (int(offs[0]) == 0 and bool(np.all(offs[1:] == offs[:-1] + lens[:-1]))), "gather_slices: DB data must be contiguous in index order"
        _CONTIGUOUS_OK.add(id(db))
    d_data = torch.empty(data.nbytes + 16, dtype=torch.uint8, device=device)
    d_offs = torch.empty(n + 1, dtype=torch.int64, device=device)
    d_lens = torch.empty(n + 1, dtype=torch.int32, device=device)
    d_keys = torch.empty(n + 1, dtype=torch.int32, device=device)
    bounds = slice_bounds(n, world)

    def byte_at(i):
        return int(offs[i]) if i < n else int(data.nbytes)

    h2d = 0
    lo, hi = bounds[rank]
    for dst, src in ((d_data[byte_at(lo): byte_at(hi)], data[byte_at(lo): byte_at(hi)]),
                     (d_offs[lo:hi], offs[lo:hi].view(np.int64)), (d_lens[lo:hi], lens[lo:hi].view(np.int32)), (d_keys[lo:hi], keys[lo:hi].view(np.int32))):
        if src.size:
            dst.copy_(torch.from_numpy(src), non_blocking=True)
            h2d += int(src.nbytes)
    for r, (a, b) in enumerate(bounds):
        if b > a:
            for t in (d_data[byte_at(a): byte_at(b)], d_offs[a:b], d_lens[a:b], d_keys[a:b]):
                dist.broadcast(t, src=r)
    return (d_data, d_offs, d_lens, d_keys), h2d


def upload_sliced(ctx, dist, db, rank, world):
    """Multi-GPU upload of a host DB (ideally pinned): PCIe carries only this rank's slice, NVLink the rest; the
    gathered arrays are adopted as the replicated device DB.  Returns (DeviceSeqDB, bytes copied host->device)."""
    import torch
    ts, h2d = gather_slices(dist, db, rank, world, torch.device("cuda", torch.cuda.current_device()))
    torch.cuda.synchronize()
    ddb = ctx.adopt(ts[0].data_ptr(), int(ts[0].numel()) - 16, ts[1].data_ptr(), ts[2].data_ptr(), ts[3].data_ptr(), int(db.n), int(db.dbtype), keepalive=ts)
    return ddb, h2d


class ShardedIteration:
    def __init__(self, ctx, dist, rank, world):
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self._t = {}
        self.last_d2h_bytes = 0
        self.exchange_bytes = 0

    def _exchange(self, counts):
        """All-to-all of the records the preceding phase left on the device; returns (recv buffer, n received)."""
        import torch
        ctx, dist = self.ctx, self.dist
        n_send = sum(counts)
        send = torch.empty(max(n_send, 1) * REC_BYTES, dtype=torch.uint8, device="cuda")
        ctx.shard_export(send.data_ptr(), n_send)
        self._ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
        self._ev[-1][0].record()
        send_counts = torch.tensor(counts, dtype=torch.int64, device="cuda")
        recv_counts = torch.empty(self.world, dtype=torch.int64, device="cuda")
        dist.all_to_all_single(recv_counts, send_counts)
        rc = [int(x) for x in recv_counts.tolist()]
        n_recv = sum(rc)
        recv = torch.empty(max(n_recv, 1) * REC_BYTES, dtype=torch.uint8, device="cuda")
        dist.all_to_all_single(recv[: n_recv * REC_BYTES], send[: n_send * REC_BYTES], split_bytes(rc), split_bytes(counts))
        self._ev[-1][1].record()
        torch.cuda.synchronize()
        self.exchange_bytes += (n_send - counts[self.rank]) * REC_BYTES
        return recv, n_recv

    def step(self, ddb, kp, rp, ep, download=False):
        import torch
        ctx = self.ctx
        self._ev, self.exchange_bytes = [], 0
        counts = ctx.shard_extract(ddb, kp, self.rank, self.world)
        recv, n_recv = self._exchange(counts)                       # all-to-all #1: k-mer records -> k-mer owner
        hist = torch.from_numpy(ctx.shard_group(ddb, kp, recv.data_ptr(), n_recv).view(np.int64)).cuda()
        del recv
        self.dist.all_reduce(hist)                                   # work per slice of the representative key space
        bounds = balanced_bounds(hist.cpu().numpy(), ddb.max_key, self.world)
        counts = ctx.shard_route(bounds)
        recv, n_recv = self._exchange(counts)                       # all-to-all #2: pair records -> representative owner
        own = (bounds[self.rank], bounds[self.rank + 1])
        self.bounds = bounds
        out, hits, alns = ctx.shard_finish(ddb, recv.data_ptr(), n_recv, own, rp, ep, want_intermediates=download)
        t2 = ctx.timings()
        t2["exchange_ms"] = sum(a.elapsed_time(b) for a, b in self._ev)
        t2["total_ms"] = t2["total_ms"] + t2["exchange_ms"]
        self._t = t2
        if download:
            host = out.download()
            self.last_d2h_bytes = int(hits.nbytes + alns.nbytes + host.data.nbytes + host.offsets.nbytes + host.lens.nbytes + host.keys.nbytes)
            del hits, alns, host
        return out

    def timings(self):
        return self._t
