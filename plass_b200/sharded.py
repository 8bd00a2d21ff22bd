"""Multi-GPU assemble iteration: one process per GPU, the sequence DB replicated in every HBM.  Each rank
extracts the k-mers of its slice of the sequences; exchange #1 routes the k-mer records to the rank that owns
the k-mer (hash of the k-mer), which sorts and groups them; exchange #2 routes the (rep, target, diagonal) pair
records to the rank that owns the representative, which finishes kmermatcher, rescorediagonal and the extension
for its queries.  The data plane is C++ (plass_b200/csrc/pg_shard.cu:
pg_shard_iteration, grouped ncclSend / ncclRecv between the stages' record buffers, DB broadcast / all-gather);
this module is the launcher-side glue: it hands the NCCL unique id from rank 0 to the others through
torch.distributed and keeps the Python mirrors of the host arithmetic for the CPU tests."""
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
    """The C++ data plane of one rank (pg_comm_init + pg_shard_iteration)."""

    def __init__(self, ctx, dist, rank, world):
        import torch
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self._t = {}
        self.last_d2h_bytes = 0
        self._phases = []
        # the NCCL unique id travels from rank 0 to the others over the launcher's own process group
        uid = ctx.comm_unique_id() if rank == 0 else bytes(api.COMM_ID_BYTES)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        ctx.comm_init(rank, world, bytes(t.cpu().numpy().tobytes()))

    def describe(self):
        return ("%d ranks, one process per GPU, C++ data plane over NCCL (pg_shard_iteration): sequence-sliced extraction, grouped ncclSend/ncclRecv of "
                "k-mer records to the k-mer owner, sort #1 + assignGroup, all-reduced work histogram -> equal-work key ranges, grouped ncclSend/ncclRecv "
                "of candidate pairs to the representative's owner, sort #2 + best diagonal + rescore + extension of the owned queries" % self.world)

    def build_and_broadcast(self, build_fn):
        """rank 0 builds the job's DB on its GPU (build_fn() -> DeviceSeqDB), every rank gets a replica over NVLink."""
        src = build_fn() if self.rank == 0 else None
        out = self.ctx.shard_broadcast_db(src, 0)
        if src is not None:
            src.free()
        return out

    def step(self, ddb, kp, rp, ep, download=False):
        """One iteration; returns this rank's slice of the new DB (DeviceSeqDB)."""
        ctx = self.ctx
        out, own, hits, alns = ctx.shard_iteration(ddb, kp, rp, ep, want_intermediates=download)
        self._t = ctx.timings()
        self.own = own
        if download:
            host = out.download()
            self.last_d2h_bytes = int(hits.nbytes + alns.nbytes + host.data.nbytes + host.offsets.nbytes + host.lens.nbytes + host.keys.nbytes)
            # the host copies live in pinned blocks of the library's pool: dropped here, so that the next step reuses them
            del hits, alns, host
        return out

    def timings(self):
        t = dict(self._t)
        ms, nb = self.ctx.shard_exchange_stats()
        t["exchange1_ms"], t["exchange2_ms"] = ms
        t["exchange1_bytes_sent"], t["exchange2_bytes_sent"] = nb
        return t

    def pinned_slice(self, ddb):
        """This rank's slice of the sequences of the replicated DB as a host DB in pinned memory (offsets rebased to 0)."""
        import torch
        host = ddb.download()
        lo, hi = slice_bounds(host.n, self.world)[self.rank]
        b0 = int(host.offsets[lo]) if lo < host.n else int(host.data.nbytes)
        b1 = int(host.offsets[hi]) if hi < host.n else int(host.data.nbytes)

        def pin(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a).view(dt)).pin_memory().numpy().view(a.dtype)
        from . import mmseqsdb
        return mmseqsdb.DB(pin(host.data[b0:b1], np.uint8), pin(host.keys[lo:hi], np.int32),
                           pin((host.offsets[lo:hi] - np.uint64(b0)).astype(np.uint64), np.int64), pin(host.lens[lo:hi], np.int32), host.dbtype)

    def upload_sliced(self, pinned):
        """PCIe carries only this rank's slice, NVLink the rest (pg_shard_allgather_db); returns the replicated DeviceSeqDB."""
        sl = self.ctx.upload(pinned)
        full = self.ctx.shard_allgather_db(sl)
        sl.free()
        return full

    def e2e_phases(self):
        return None

    def verify_against_single_gpu(self, ddb, kp, rp, ep):
        """The sharded iteration's results (hits, alignments, all-gathered new DB) against a single-GPU run of the same
        replicated DB on rank 0.  Counts and order-independent checksums of the hit / alignment records are summed over
        the ranks; the all-gathered DB (the next iteration's input) must equal the single-GPU output byte for byte."""
        import torch
        ctx, dist = self.ctx, self.dist
        out, own, hits, alns = ctx.shard_iteration(ddb, kp, rp, ep, want_intermediates=True)
        full = ctx.shard_allgather_db(out)
        out.free()
        mine = np.array([len(hits), len(alns), record_checksum(hits), record_checksum(alns)], dtype=np.uint64)
        t = torch.from_numpy(mine.view(np.int64)).cuda()
        dist.all_reduce(t)                       # int64 sums wrap like uint64 sums
        tot = t.cpu().numpy().view(np.uint64)
        del hits, alns
        ctx.shard_release_buffers()              # collective: rank 0 needs the memory for the whole job on one GPU
        res = None
        if self.rank == 0:
            g = full.download()
            full.free()
            one, h1, a1 = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
            o = one.download()
            one.free()
            ctx.release_workspace()              # the whole job's record buffers: the sharded steps that follow need the room
            want = np.array([len(h1), len(a1), record_checksum(h1), record_checksum(a1)], dtype=np.uint64)
            eq = {f: bool(np.array_equal(getattr(g, f), getattr(o, f))) for f in ("keys", "lens", "offsets", "data")}
            same_db = all(eq.values())
            detail = None
            if not same_db:
                detail = {"arrays_equal": eq, "n": [int(g.n), int(o.n)], "data_bytes": [int(g.data.nbytes), int(o.data.nbytes)]}
                if g.n == o.n:
                    bad_len = np.flatnonzero(np.asarray(g.lens) != np.asarray(o.lens))
                    detail["entries_with_other_length"] = int(len(bad_len))
                    if eq["lens"] and eq["offsets"] and g.data.nbytes == o.data.nbytes:
                        diff = np.flatnonzero(np.asarray(g.data) != np.asarray(o.data))
                        detail["differing_bytes"] = int(len(diff))
                        if len(diff):
                            ent = np.unique(np.searchsorted(np.asarray(o.offsets), diff, side="right") - 1)
                            detail["differing_entries"] = int(len(ent))
                            detail["first_entries"] = [{"index": int(i), "key": int(o.keys[i]), "sharded": g.entry(int(i)).decode("latin1")[:120],
                                                        "single": o.entry(int(i)).decode("latin1")[:120]} for i in ent[:4]]
                    elif len(bad_len):
                        detail["first_entries"] = [{"index": int(i), "key": int(o.keys[i]), "sharded": g.entry(int(i)).decode("latin1")[:120],
                                                    "single": o.entry(int(i)).decode("latin1")[:120]} for i in bad_len[:4]]
            res = {"ranks": self.world, "hits": int(tot[0]), "hits_single_gpu": int(want[0]), "alignments": int(tot[1]), "alignments_single_gpu": int(want[1]),
                   "hit_checksum_equal": bool(tot[2] == want[2]), "alignment_checksum_equal": bool(tot[3] == want[3]),
                   "gathered_db_equals_single_gpu_db": bool(same_db), "sequences": int(o.n), "detail": detail}
            res["equal"] = bool(tot[0] == want[0] and tot[1] == want[1] and res["hit_checksum_equal"] and res["alignment_checksum_equal"] and same_db)
        else:
            full.free()
        dist.barrier()
        return res


def record_checksum(recs):
    """Order-independent checksum of an array of POD records: sum over the records of a mixed hash of their 64-bit words
    (wrap-around uint64 arithmetic)."""
    if len(recs) == 0:
        return np.uint64(0)
    w = np.ascontiguousarray(recs).view(np.uint64).reshape(len(recs), -1)
    with np.errstate(over="ignore"):
        h = np.zeros(len(recs), dtype=np.uint64)
        for j in range(w.shape[1]):
            h = (h ^ w[:, j]) * np.uint64(0x9E3779B97F4A7C15)
            h ^= h >> np.uint64(29)
        return np.uint64(h.sum(dtype=np.uint64))
