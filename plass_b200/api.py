"""ctypes binding of libplassgpu.so -- the host-side mirror of the C ABI in include/plassgpu.h.

Function names and argument meaning follow the reference's three hot-path commands
(kmermatcher / rescorediagonal / assembleresults); see INTEGRATION.md.  There is no CPU fallback:
loading works without a GPU (so that the symbol table can be checked), every compute call fails
loudly when no CUDA device is present.
"""
import ctypes as C
import os
import weakref
import numpy as np

from . import mmseqsdb

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplassgpu.so")

# symbols include/plassgpu.h declares
EXPORTS = ["pg_last_error", "pg_device_count", "pg_init", "pg_destroy", "pg_get_timings", "pg_seqdb_upload", "pg_seqdb_adopt",
           "pg_seqdb_download", "pg_seqdb_size", "pg_seqdb_free", "pg_kmermatch", "pg_rescore", "pg_extend",
           "pg_assemble_iteration", "pg_free_host", "pg_set_async_results", "pg_results_ticket", "pg_results_wait",
           "pg_shard_pairs", "pg_shard_extract", "pg_shard_group", "pg_shard_route", "pg_shard_export", "pg_shard_finish", "pg_shard_owner_range", "pg_seqdb_max_key",
           "pg_seqdb_upload_async", "pg_findassemblystart", "pg_assemble_step0", "pg_cyclecheck", "pg_extractorfs", "pg_translatenucs",
           "pg_seqdb_concat", "pg_set_split_memory_limit", "pg_comm_unique_id", "pg_comm_init", "pg_comm_destroy", "pg_comm_rank", "pg_comm_world",
           "pg_shard_broadcast_db", "pg_shard_allgather_db", "pg_shard_iteration", "pg_shard_exchange_stats", "pg_shard_balanced_bounds", "pg_shard_release_buffers", "pg_release_workspace"]


SHARD_HIST_BINS = 4096   # PG_SHARD_HIST_BINS
COMM_ID_BYTES = 128      # PG_COMM_ID_BYTES


class SeqDBView(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_bytes", C.c_uint64), ("offsets", C.c_void_p), ("lens", C.c_void_p),
                ("keys", C.c_void_p), ("n", C.c_uint64), ("dbtype", C.c_int)]


class KmParams(C.Structure):
    _fields_ = [("kmer_size", C.c_int), ("alph_size", C.c_int), ("kmers_per_seq", C.c_int),
                ("kmers_per_seq_scale", C.c_float), ("hash_shift", C.c_int), ("include_only_extendable", C.c_int),
                ("ignore_multi_kmer", C.c_int), ("cov_mode", C.c_int), ("cov_thr", C.c_float),
                ("hash_start", C.c_uint32), ("hash_end", C.c_uint32)]


class RsParams(C.Structure):
    _fields_ = [("rescore_mode", C.c_int), ("seq_id_thr", C.c_float), ("eval_thr", C.c_double), ("cov_mode", C.c_int),
                ("cov_thr", C.c_float), ("aln_len_thr", C.c_int), ("seq_id_mode", C.c_int)]


class ExParams(C.Structure):
    _fields_ = [("seq_id_thr", C.c_float), ("max_seq_len", C.c_int), ("keep_target", C.c_int), ("rescore_mode", C.c_int)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("extract_ms", "sort1_ms", "group_ms", "sort2_ms", "reduce_ms", "rescore_ms",
                                         "extend_ms", "exchange_ms", "total_ms")] + \
               [(n, C.c_uint64) for n in ("n_kmer_records", "n_pair_records", "n_hits", "n_alns", "n_extended",
                                          "kernel_launches", "sort1_bytes")] + \
               [("sort1_scatter_ms", C.c_float), ("sort1_passes", C.c_uint32), ("splits", C.c_uint32), ("spilled_records", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


HIT = np.dtype([("rep", "<u4"), ("target", "<u4"), ("score", "<i4"), ("diag", "<i4")], align=True)
ALN = np.dtype([("query", "<u4"), ("target", "<u4"), ("bits", "<i4"), ("seq_id", "<f4"), ("evalue", "<f8"),
                ("q_start", "<i4"), ("q_end", "<i4"), ("q_len", "<i4"),
                ("db_start", "<i4"), ("db_end", "<i4"), ("db_len", "<i4")], align=True)


class OrfParams(C.Structure):
    """extractorfs flags (include/plassgpu.h pg_orf_params)."""
    _fields_ = [("min_length", C.c_int), ("max_length", C.c_int), ("max_gaps", C.c_int), ("contig_start_mode", C.c_int),
                ("contig_end_mode", C.c_int), ("orf_start_mode", C.c_int), ("forward_frames", C.c_uint), ("reverse_frames", C.c_uint),
                ("translation_table", C.c_int), ("use_all_table_starts", C.c_int)]


def orf_params_long(min_length=45):
    """EXTRACTORFS_LONG_PAR of the assemble workflow (Assembler.cpp:114-118)."""
    return OrfParams(min_length=min_length, max_length=32734, max_gaps=0, contig_start_mode=2, contig_end_mode=2, orf_start_mode=0,
                     forward_frames=7, reverse_frames=7, translation_table=1, use_all_table_starts=0)


def orf_params_start(min_length=45):
    """EXTRACTORFS_START_PAR (Assembler.cpp:121-128): fragments that begin at an ATG and run off the read end."""
    return OrfParams(min_length=min(min_length, 20), max_length=min_length, max_gaps=0, contig_start_mode=1, contig_end_mode=0, orf_start_mode=0,
                     forward_frames=7, reverse_frames=7, translation_table=1, use_all_table_starts=0)


def default_km_params(nucl=False, **kw):
    """Workflow defaults: Assembler.cpp:10-27 (aa) / Nuclassembler.cpp:10-32 (nt)."""
    p = KmParams(kmer_size=22 if nucl else 14, alph_size=5 if nucl else 13, kmers_per_seq=60,
                 kmers_per_seq_scale=0.1 if nucl else 0.0, hash_shift=67, include_only_extendable=1 if nucl else 0,
                 ignore_multi_kmer=1, cov_mode=0, cov_thr=0.0, hash_start=0, hash_end=65535)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_rs_params(nucl=False, **kw):
    p = RsParams(rescore_mode=3, seq_id_thr=0.99 if nucl else 0.9, eval_thr=1e-5, cov_mode=0, cov_thr=0.0, aln_len_thr=0, seq_id_mode=0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_ex_params(nucl=False, **kw):
    p = ExParams(seq_id_thr=0.99 if nucl else 0.9, max_seq_len=200000 if nucl else 65535, keep_target=1, rescore_mode=3)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


_lib = None


def load_library():
    """dlopen the in-tree CUDA library.  Raises if it has not been built (python -m plass_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libplassgpu.so is missing -- build it with `python -m plass_b200.build`; there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.pg_last_error.restype = C.c_char_p
        _lib.pg_seqdb_size.restype = C.c_uint64
        _lib.pg_seqdb_size.argtypes = [C.c_void_p]
        _lib.pg_free_host.argtypes = [C.c_void_p]
        _lib.pg_destroy.argtypes = [C.c_void_p]
        _lib.pg_seqdb_free.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


class PlassGpuError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise PlassGpuError("%s failed: %s" % (what, load_library().pg_last_error().decode()))


def _take(ptr, n, dtype):
    """Wrap a pinned host array returned by the library as a numpy array WITHOUT copying; the block goes
    back to the library's pinned pool (pg_free_host) when the array is garbage collected."""
    lib = load_library()
    if n == 0 or not ptr.value:
        if ptr.value:
            lib.pg_free_host(ptr)
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
    out = np.frombuffer(buf, dtype=dtype)
    weakref.finalize(buf, lib.pg_free_host, C.c_void_p(ptr.value))
    return out


class DeviceSeqDB:
    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.handle = handle
        self.keepalive = None

    @property
    def n(self):
        return int(load_library().pg_seqdb_size(self.handle))

    def download(self):
        lib = load_library()
        d, o, l, k = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nb, n = C.c_uint64(), C.c_uint64()
        _check(lib.pg_seqdb_download(self.ctx.handle, self.handle, C.byref(d), C.byref(nb), C.byref(o), C.byref(l), C.byref(k), C.byref(n)), "pg_seqdb_download")
        return mmseqsdb.DB(_take(d, nb.value, np.dtype("u1")), _take(k, n.value, np.dtype("<u4")),
                           _take(o, n.value, np.dtype("<u8")), _take(l, n.value, np.dtype("<u4")), self.dbtype)

    @property
    def max_key(self):
        lib = load_library()
        lib.pg_seqdb_max_key.restype = C.c_uint32
        lib.pg_seqdb_max_key.argtypes = [C.c_void_p]
        return int(lib.pg_seqdb_max_key(self.handle))

    def free(self):
        if self.handle:
            load_library().pg_seqdb_free(self.ctx.handle, self.handle)
            self.handle = None
            self.keepalive = None


class Context:
    """One GPU.  Methods mirror the reference commands they replace."""

    def __init__(self, device=0):
        lib = load_library()
        h = C.c_void_p()
        _check(lib.pg_init(C.c_int(device), C.byref(h)), "pg_init")
        self.handle = h
        self.device = device

    def close(self):
        if self.handle:
            load_library().pg_destroy(self.handle)
            self.handle = None

    # asynchronous results: assemble_iteration(want_intermediates=True) and DeviceSeqDB.download() only enqueue their
    # device -> host copies; the returned arrays may be read after results_wait(t) for a ticket t taken after the calls
    def set_async_results(self, on):
        _check(load_library().pg_set_async_results(self.handle, C.c_int(1 if on else 0)), "pg_set_async_results")

    def results_ticket(self):
        t = C.c_uint64()
        _check(load_library().pg_results_ticket(self.handle, C.byref(t)), "pg_results_ticket")
        return int(t.value)

    def results_wait(self, ticket):
        _check(load_library().pg_results_wait(self.handle, C.c_uint64(ticket)), "pg_results_wait")

    def set_split_memory_limit(self, nbytes):
        """--split-memory-limit: bound on the kmermatcher stage's two record buffers (0 = 90 % of the free device memory)."""
        _check(load_library().pg_set_split_memory_limit(self.handle, C.c_uint64(int(nbytes))), "pg_set_split_memory_limit")

    def debug_force_splits(self, n):
        _check(load_library().pg_debug_force_splits(self.handle, C.c_uint(int(n))), "pg_debug_force_splits")

    def release_workspace(self):
        _check(load_library().pg_release_workspace(self.handle), "pg_release_workspace")

    def timings(self):
        t = Timings()
        _check(load_library().pg_get_timings(self.handle, C.byref(t)), "pg_get_timings")
        return t.as_dict()

    def upload(self, db):
        """db: mmseqsdb.DB (host).  Returns a DeviceSeqDB resident in HBM."""
        v = SeqDBView()
        data = np.ascontiguousarray(db.data)
        offs = np.ascontiguousarray(db.offsets, dtype=np.uint64)
        lens = np.ascontiguousarray(db.lens, dtype=np.uint32)
        keys = np.ascontiguousarray(db.keys, dtype=np.uint32)
        v.data, v.data_bytes = data.ctypes.data, data.nbytes
        v.offsets, v.lens, v.keys, v.n, v.dbtype = offs.ctypes.data, lens.ctypes.data, keys.ctypes.data, db.n, db.dbtype
        h = C.c_void_p()
        _check(load_library().pg_seqdb_upload(self.handle, C.byref(v), C.byref(h)), "pg_seqdb_upload")
        d = DeviceSeqDB(self, h)
        d.dbtype = db.dbtype
        return d

    def upload_async(self, db):
        """Like upload(), but only enqueues the host -> device copies (on the context's upload stream) and returns: the
        transfer runs underneath whatever the GPU is computing.  db's arrays must be pinned and are kept alive by the
        returned object."""
        v = SeqDBView()
        data = np.ascontiguousarray(db.data)
        offs = np.ascontiguousarray(db.offsets, dtype=np.uint64)
        lens = np.ascontiguousarray(db.lens, dtype=np.uint32)
        keys = np.ascontiguousarray(db.keys, dtype=np.uint32)
        v.data, v.data_bytes = data.ctypes.data, data.nbytes
        v.offsets, v.lens, v.keys, v.n, v.dbtype = offs.ctypes.data, lens.ctypes.data, keys.ctypes.data, db.n, db.dbtype
        h = C.c_void_p()
        _check(load_library().pg_seqdb_upload_async(self.handle, C.byref(v), C.byref(h)), "pg_seqdb_upload_async")
        d = DeviceSeqDB(self, h)
        d.dbtype = db.dbtype
        d.keepalive = (data, offs, lens, keys)
        return d

    def adopt(self, data_ptr, data_bytes, offsets_ptr, lens_ptr, keys_ptr, n, dbtype, keepalive=None):
        """Wraps caller-owned device arrays (e.g. torch tensors filled by an all-gather) as a DeviceSeqDB."""
        v = SeqDBView()
        v.data, v.data_bytes = data_ptr, data_bytes
        v.offsets, v.lens, v.keys, v.n, v.dbtype = offsets_ptr, lens_ptr, keys_ptr, n, dbtype
        h = C.c_void_p()
        _check(load_library().pg_seqdb_adopt(self.handle, C.byref(v), C.byref(h)), "pg_seqdb_adopt")
        d = DeviceSeqDB(self, h)
        d.dbtype = dbtype
        d.keepalive = keepalive
        return d

    # kmermatcher (lib/mmseqs/src/linclust/kmermatcher.cpp:780)
    def kmermatcher(self, ddb, kp):
        out, n = C.c_void_p(), C.c_uint64()
        _check(load_library().pg_kmermatch(self.handle, ddb.handle, C.byref(kp), C.byref(out), C.byref(n)), "pg_kmermatch")
        return _take(out, n.value, HIT)

    # rescorediagonal (lib/mmseqs/src/alignment/rescorediagonal.cpp:381)
    def rescorediagonal(self, ddb, hits, rp):
        hits = np.ascontiguousarray(hits, dtype=HIT)
        out, n = C.c_void_p(), C.c_uint64()
        _check(load_library().pg_rescore(self.handle, ddb.handle, C.c_void_p(hits.ctypes.data), C.c_uint64(len(hits)),
                                         C.byref(rp), C.byref(out), C.byref(n)), "pg_rescore")
        return _take(out, n.value, ALN)

    # assembleresults / nuclassembleresults (src/assembler/assembleresult.cpp:358, nuclassembleresult.cpp:400)
    def assembleresults(self, ddb, alns, ep):
        alns = np.ascontiguousarray(alns, dtype=ALN)
        h, ext = C.c_void_p(), C.c_void_p()
        _check(load_library().pg_extend(self.handle, ddb.handle, C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)),
                                        C.byref(ep), C.byref(h), C.byref(ext)), "pg_extend")
        out = DeviceSeqDB(self, h)
        out.dbtype = ddb.dbtype
        return out, _take(ext, out.n, np.dtype("u1"))

    # findassemblystart (src/assembler/findassemblystart.cpp:35-176)
    def findassemblystart(self, ddb, alns):
        """Returns (corrected DeviceSeqDB, add_stop int32 per sequence: cut position or -1)."""
        alns = np.ascontiguousarray(alns, dtype=ALN)
        h, st = C.c_void_p(), C.c_void_p()
        _check(load_library().pg_findassemblystart(self.handle, ddb.handle, C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)),
                                                   C.byref(h), C.byref(st)), "pg_findassemblystart")
        out = DeviceSeqDB(self, h)
        out.dbtype = ddb.dbtype
        return out, _take(st, ddb.n, np.dtype("<i4"))

    def assemble_step0(self, ddb, kp, rp, ep, want_intermediates=False):
        """plass STEP 0 fused in HBM.  Returns (corrected DeviceSeqDB, assembly_0 DeviceSeqDB, hits or None, alns or None);
        hits / alns are those of the corrected pass (pref_corrected_0 / aln_corrected_0)."""
        hc, ho = C.c_void_p(), C.c_void_p()
        if want_intermediates:
            hp, hn, ap, an = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
            _check(load_library().pg_assemble_step0(self.handle, ddb.handle, C.byref(kp), C.byref(rp), C.byref(ep), C.byref(hc), C.byref(ho),
                                                    C.byref(hp), C.byref(hn), C.byref(ap), C.byref(an)), "pg_assemble_step0")
            hits, alns = _take(hp, hn.value, HIT), _take(ap, an.value, ALN)
        else:
            _check(load_library().pg_assemble_step0(self.handle, ddb.handle, C.byref(kp), C.byref(rp), C.byref(ep), C.byref(hc), C.byref(ho),
                                                    None, None, None, None), "pg_assemble_step0")
            hits = alns = None
        corr, out = DeviceSeqDB(self, hc), DeviceSeqDB(self, ho)
        corr.dbtype = out.dbtype = ddb.dbtype
        return corr, out, hits, alns

    # extractorfs [+ translatenucs --add-orf-stop 1] (lib/mmseqs/src/util/extractorfs.cpp:20, translatenucs.cpp:14)
    def extractorfs(self, ddb, op, translate=False, want_info=True):
        """Returns (fragment DeviceSeqDB keyed 0..n-1, orf_info uint32 (n, 4): read key, fromPos, toPos, flags -- or None)."""
        h, info = C.c_void_p(), C.c_void_p()
        _check(load_library().pg_extractorfs(self.handle, ddb.handle, C.byref(op), C.c_int(1 if translate else 0), C.byref(h),
                                             C.byref(info) if want_info else None), "pg_extractorfs")
        out = DeviceSeqDB(self, h)
        out.dbtype = 0 if translate else 1
        n = out.n
        return out, (_take(info, 4 * n, np.dtype("<u4")).reshape(n, 4) if want_info else None)

    def translatenucs(self, ddb, flags=None, translation_table=1):
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        h = C.c_void_p()
        _check(load_library().pg_translatenucs(self.handle, ddb.handle, None if fl is None else C.c_void_p(fl.ctypes.data), C.c_int(translation_table),
                                               C.byref(h)), "pg_translatenucs")
        out = DeviceSeqDB(self, h)
        out.dbtype = 0
        return out

    # concatdbs (two sequence DBs, keys renumbered)
    def concat(self, a, b):
        h = C.c_void_p()
        _check(load_library().pg_seqdb_concat(self.handle, a.handle, b.handle, C.byref(h)), "pg_seqdb_concat")
        out = DeviceSeqDB(self, h)
        out.dbtype = a.dbtype
        return out

    def six_frame_fragments(self, reads_ddb, min_length=45):
        """nucl_reads -> aa_6f_start_long as data/assemble.sh:41-77 builds it: long ORFs first, then the start fragments."""
        lo, _ = self.extractorfs(reads_ddb, orf_params_long(min_length), translate=True, want_info=False)
        st, _ = self.extractorfs(reads_ddb, orf_params_start(min_length), translate=True, want_info=False)
        out = self.concat(lo, st)
        lo.free(); st.free()
        return out

    # cyclecheck (src/assembler/cyclecheck.cpp:71-274)
    def cyclecheck(self, ddb, max_seq_len):
        """Split diagonal per sequence (uint32, 0 = not circular)."""
        sp = C.c_void_p()
        _check(load_library().pg_cyclecheck(self.handle, ddb.handle, C.c_int(int(max_seq_len)), C.byref(sp)), "pg_cyclecheck")
        return _take(sp, ddb.n, np.dtype("<u4"))

    def assemble_iteration(self, ddb, kp, rp, ep, want_intermediates=False):
        """One fused iteration in HBM.  Returns (next DeviceSeqDB, hits or None, alns or None)."""
        h = C.c_void_p()
        if want_intermediates:
            ho, hn, ao, an = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
            _check(load_library().pg_assemble_iteration(self.handle, ddb.handle, C.byref(kp), C.byref(rp), C.byref(ep), C.byref(h),
                                                        C.byref(ho), C.byref(hn), C.byref(ao), C.byref(an)), "pg_assemble_iteration")
            hits, alns = _take(ho, hn.value, HIT), _take(ao, an.value, ALN)
        else:
            _check(load_library().pg_assemble_iteration(self.handle, ddb.handle, C.byref(kp), C.byref(rp), C.byref(ep), C.byref(h),
                                                        None, None, None, None), "pg_assemble_iteration")
            hits = alns = None
        out = DeviceSeqDB(self, h)
        out.dbtype = ddb.dbtype
        return out, hits, alns

    # ---- multi-GPU data plane in C++ over NCCL (plass_b200/csrc/pg_shard.cu) ----
    @staticmethod
    def _torch_nccl_first():
        """libplassgpu.so loads NCCL at run time by soname.  In a Python process that will also import PyTorch, PyTorch's own
        (newer) libnccl.so.2 has to be the copy in the process: import torch before the first NCCL call if it is installed."""
        try:
            import torch  # noqa: F401
        except ImportError:
            pass

    @staticmethod
    def comm_unique_id():
        Context._torch_nccl_first()
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        _check(load_library().pg_comm_unique_id(buf), "pg_comm_unique_id")
        return bytes(buf)

    def comm_init(self, rank, world, comm_id):
        Context._torch_nccl_first()
        buf = (C.c_uint8 * COMM_ID_BYTES)(*comm_id)
        _check(load_library().pg_comm_init(self.handle, C.c_int(rank), C.c_int(world), buf), "pg_comm_init")

    def shard_broadcast_db(self, ddb, root, dbtype=0):
        h = C.c_void_p()
        _check(load_library().pg_shard_broadcast_db(self.handle, ddb.handle if ddb is not None else None, C.c_int(root), C.byref(h)), "pg_shard_broadcast_db")
        out = DeviceSeqDB(self, h)
        out.dbtype = ddb.dbtype if ddb is not None else dbtype
        return out

    def shard_allgather_db(self, slice_db):
        h = C.c_void_p()
        _check(load_library().pg_shard_allgather_db(self.handle, slice_db.handle, C.byref(h)), "pg_shard_allgather_db")
        out = DeviceSeqDB(self, h)
        out.dbtype = slice_db.dbtype
        return out

    def shard_iteration(self, ddb, kp, rp, ep, want_intermediates=False):
        """One iteration over all ranks of the communicator.  Returns (DeviceSeqDB of the owned keys, (own_lo, own_hi), hits or
        None, alns or None) -- this rank's share."""
        lib = load_library()
        h, lo, hi = C.c_void_p(), C.c_uint32(), C.c_uint32()
        if want_intermediates:
            ho, hn, ao, an = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
            _check(lib.pg_shard_iteration(self.handle, ddb.handle, C.byref(kp), C.byref(rp), C.byref(ep), C.byref(h), C.byref(lo), C.byref(hi),
                                          C.byref(ho), C.byref(hn), C.byref(ao), C.byref(an)), "pg_shard_iteration")
            hits, alns = _take(ho, hn.value, HIT), _take(ao, an.value, ALN)
        else:
            _check(lib.pg_shard_iteration(self.handle, ddb.handle, C.byref(kp), C.byref(rp), C.byref(ep), C.byref(h), C.byref(lo), C.byref(hi),
                                          None, None, None, None), "pg_shard_iteration")
            hits = alns = None
        out = DeviceSeqDB(self, h)
        out.dbtype = ddb.dbtype
        return out, (int(lo.value), int(hi.value)), hits, alns

    def shard_release_buffers(self):
        _check(load_library().pg_shard_release_buffers(self.handle), "pg_shard_release_buffers")

    def shard_exchange_stats(self):
        ms, nb = (C.c_float * 2)(), (C.c_uint64 * 2)()
        _check(load_library().pg_shard_exchange_stats(self.handle, ms, nb), "pg_shard_exchange_stats")
        return [float(x) for x in ms], [int(x) for x in nb]

    # ---- multi-GPU phases, one call each (the exchanges are the caller's; see tests/test_gpu_parity.py) ----
    def shard_pairs(self, ddb, kp, world):
        counts = (C.c_uint64 * world)()
        _check(load_library().pg_shard_pairs(self.handle, ddb.handle, C.byref(kp), C.c_int(world), counts), "pg_shard_pairs")
        return [int(c) for c in counts]

    def shard_extract(self, ddb, kp, rank, world):
        counts = (C.c_uint64 * world)()
        _check(load_library().pg_shard_extract(self.handle, ddb.handle, C.byref(kp), C.c_int(rank), C.c_int(world), counts), "pg_shard_extract")
        return [int(x) for x in counts]

    def shard_group(self, ddb, kp, device_ptr, n_records):
        hist = np.zeros(SHARD_HIST_BINS, dtype=np.uint64)
        _check(load_library().pg_shard_group(self.handle, ddb.handle, C.byref(kp), C.c_void_p(device_ptr), C.c_uint64(n_records),
                                             C.c_void_p(hist.ctypes.data)), "pg_shard_group")
        return hist

    def shard_route(self, bounds):
        world = len(bounds) - 1
        b = (C.c_uint32 * (world + 1))(*[int(x) for x in bounds])
        counts = (C.c_uint64 * world)()
        _check(load_library().pg_shard_route(self.handle, C.c_int(world), b, counts), "pg_shard_route")
        return [int(x) for x in counts]

    def shard_export(self, device_ptr, n_records):
        _check(load_library().pg_shard_export(self.handle, C.c_void_p(device_ptr), C.c_uint64(n_records)), "pg_shard_export")

    def shard_finish(self, ddb, device_ptr, n_pairs, own, rp, ep, want_intermediates=False):
        h = C.c_void_p()
        lib = load_library()
        if want_intermediates:
            ho, hn, ao, an = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
            _check(lib.pg_shard_finish(self.handle, ddb.handle, C.c_void_p(device_ptr), C.c_uint64(n_pairs), C.c_uint32(own[0]), C.c_uint32(own[1]),
                                       C.byref(rp), C.byref(ep), C.byref(h), C.byref(ho), C.byref(hn), C.byref(ao), C.byref(an)), "pg_shard_finish")
            hits, alns = _take(ho, hn.value, HIT), _take(ao, an.value, ALN)
        else:
            _check(lib.pg_shard_finish(self.handle, ddb.handle, C.c_void_p(device_ptr), C.c_uint64(n_pairs), C.c_uint32(own[0]), C.c_uint32(own[1]),
                                       C.byref(rp), C.byref(ep), C.byref(h), None, None, None, None), "pg_shard_finish")
            hits = alns = None
        out = DeviceSeqDB(self, h)
        out.dbtype = ddb.dbtype
        return out, hits, alns

    # diagnostics
    def debug_extract(self, ddb, kp):
        out, n = C.c_void_p(), C.c_uint64()
        _check(load_library().pg_debug_extract(self.handle, ddb.handle, C.byref(kp), C.byref(out), C.byref(n)), "pg_debug_extract")
        return _take(out, n.value, np.dtype([("w0", "<u8"), ("w1", "<u8")]))

    def debug_radix_sort(self, recs, ranges):
        """recs: (n,2) uint64; ranges: list of (word, lo, hi) least-significant first."""
        recs = np.ascontiguousarray(recs, dtype=np.uint64)
        w = (C.c_int * len(ranges))(*[r[0] for r in ranges])
        lo = (C.c_int * len(ranges))(*[r[1] for r in ranges])
        hi = (C.c_int * len(ranges))(*[r[2] for r in ranges])
        _check(load_library().pg_debug_radix_sort(self.handle, C.c_void_p(recs.ctypes.data), C.c_uint64(len(recs)), w, lo, hi, len(ranges)), "pg_debug_radix_sort")
        return recs
