"""ctypes binding of oracle/liboracle.so (the CPU oracle -- test infrastructure only)."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

U64MAX = (1 << 64) - 1


class SeqDB(C.Structure):
    _fields_ = [("data", C.c_void_p), ("offsets", C.c_void_p), ("lens", C.c_void_p), ("keys", C.c_void_p),
                ("n", C.c_uint64), ("dbtype", C.c_int)]


class KmParams(C.Structure):
    _fields_ = [("kmer_size", C.c_int), ("alph_size", C.c_int), ("kmers_per_seq", C.c_int),
                ("kmers_per_seq_scale", C.c_float), ("hash_shift", C.c_int), ("include_only_extendable", C.c_int),
                ("ignore_multi_kmer", C.c_int), ("cov_mode", C.c_int), ("cov_thr", C.c_float),
                ("hash_start", C.c_uint64), ("hash_end", C.c_uint64)]


class RsParams(C.Structure):
    _fields_ = [("rescore_mode", C.c_int), ("seq_id_thr", C.c_float), ("eval_thr", C.c_double), ("cov_mode", C.c_int),
                ("cov_thr", C.c_float), ("aln_len_thr", C.c_int), ("seq_id_mode", C.c_int)]


class ExParams(C.Structure):
    _fields_ = [("seq_id_thr", C.c_float), ("max_seq_len", C.c_int), ("keep_target", C.c_int), ("rescore_mode", C.c_int)]


KMER_REC = np.dtype([("kmer", "<u8"), ("id", "<u4"), ("seq_len", "<i4"), ("pos", "<i4")], align=True)
HIT = np.dtype([("rep", "<u4"), ("target", "<u4"), ("score", "<i4"), ("diag", "<i4")], align=True)
ALN = np.dtype([("query", "<u4"), ("target", "<u4"), ("bits", "<i4"), ("seq_id", "<f4"), ("evalue", "<f8"),
                ("q_start", "<i4"), ("q_end", "<i4"), ("q_len", "<i4"),
                ("db_start", "<i4"), ("db_end", "<i4"), ("db_len", "<i4")], align=True)
assert KMER_REC.itemsize == 24 and HIT.itemsize == 16 and ALN.itemsize == 48


def build():
    srcs = [os.path.join(ORACLE_DIR, "oracle.cpp"), os.path.join(ORACLE_DIR, "oracle_next.cpp")]
    deps = srcs + [os.path.join(ORACLE_DIR, "oracle.h"), os.path.join(ORACLE_DIR, "oracle_tables.h")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB_PATH] + srcs, check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.or_hash_u64.restype = C.c_uint64
        _lib.or_hash_u64.argtypes = [C.c_uint64, C.c_uint64]
        for f in ("or_evalue", "or_bitscore", "or_raw_from_bits"):
            getattr(_lib, f).restype = C.c_double
        _lib.or_evalue.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        _lib.or_bitscore.argtypes = [C.c_int, C.c_double]
        _lib.or_raw_from_bits.argtypes = [C.c_int, C.c_double]
        _lib.or_free.argtypes = [C.c_void_p]
    return _lib


def seqdb_struct(db):
    """db: plass_b200.mmseqsdb.DB (arrays must stay alive while the struct is used)."""
    s = SeqDB()
    s.data = db.data.ctypes.data
    s.offsets = db.offsets.ctypes.data
    s.lens = db.lens.ctypes.data
    s.keys = db.keys.ctypes.data
    s.n = db.n
    s.dbtype = db.dbtype
    return s


def _take(ptr, n, dtype):
    if n == 0:
        out = np.zeros(0, dtype=dtype)
    else:
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=dtype).copy()
    lib().or_free(ptr)
    return out


def extract_kmers(db, kp):
    out, n = C.c_void_p(), C.c_uint64()
    s = seqdb_struct(db)
    rc = lib().or_extract_kmers(C.byref(s), C.byref(kp), C.byref(out), C.byref(n))
    assert rc == 0
    return _take(out, n.value, KMER_REC)


def kmermatch(db, kp):
    out, n = C.c_void_p(), C.c_uint64()
    s = seqdb_struct(db)
    rc = lib().or_kmermatch(C.byref(s), C.byref(kp), C.byref(out), C.byref(n))
    assert rc == 0
    return _take(out, n.value, HIT)


def rescore(db, hits, rp):
    out, n = C.c_void_p(), C.c_uint64()
    s = seqdb_struct(db)
    hits = np.ascontiguousarray(hits)
    rc = lib().or_rescore(C.byref(s), C.c_void_p(hits.ctypes.data), C.c_uint64(len(hits)), C.byref(rp), C.byref(out), C.byref(n))
    assert rc == 0, rc
    return _take(out, n.value, ALN)


def extend(db, alns, ep):
    from plass_b200.mmseqsdb import DB
    od, oo, ol, ok, ex = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    on, ob = C.c_uint64(), C.c_uint64()
    s = seqdb_struct(db)
    alns = np.ascontiguousarray(alns)
    rc = lib().or_extend(C.byref(s), C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)), C.byref(ep),
                         C.byref(od), C.byref(oo), C.byref(ol), C.byref(ok), C.byref(ex), C.byref(on), C.byref(ob))
    assert rc == 0, rc
    n = on.value
    data = _take(od, ob.value, np.dtype("u1"))
    offs = _take(oo, n, np.dtype("<u8"))
    lens = _take(ol, n, np.dtype("<u4"))
    keys = _take(ok, n, np.dtype("<u4"))
    ext = _take(ex, n, np.dtype("u1"))
    return DB(data, keys, offs, lens, db.dbtype), ext


def findstart(db, alns):
    """findassemblystart: returns (new DB, add_stop[int32 per sequence])."""
    from plass_b200.mmseqsdb import DB
    od, oo, ol, ok, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    on, ob = C.c_uint64(), C.c_uint64()
    s = seqdb_struct(db)
    alns = np.ascontiguousarray(alns)
    rc = lib().or_findstart(C.byref(s), C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)),
                            C.byref(od), C.byref(oo), C.byref(ol), C.byref(ok), C.byref(on), C.byref(ob), C.byref(st))
    assert rc == 0, rc
    n = on.value
    out = DB(_take(od, ob.value, np.dtype("u1")), _take(ok, n, np.dtype("<u4")), _take(oo, n, np.dtype("<u8")),
             _take(ol, n, np.dtype("<u4")), db.dbtype)
    return out, _take(st, db.n, np.dtype("<i4"))


class OrfParams(C.Structure):
    _fields_ = [("min_length", C.c_int), ("max_length", C.c_int), ("max_gaps", C.c_int), ("contig_start_mode", C.c_int),
                ("contig_end_mode", C.c_int), ("orf_start_mode", C.c_int), ("forward_frames", C.c_uint), ("reverse_frames", C.c_uint),
                ("translation_table", C.c_int), ("use_all_table_starts", C.c_int)]


def orf_params_from_flags(flags, cls=None):
    """extractorfs argv flags (dict) -> OrfParams (or the GPU library's identical struct)."""
    def frames(s):
        m = 0
        for x in s.split(","):
            m |= {"1": 1, "2": 2, "3": 4}.get(x, 0)
        return m
    cls = cls or OrfParams
    return cls(min_length=int(flags.get("--min-length", 30)), max_length=int(flags.get("--max-length", 32734)),
               max_gaps=int(flags.get("--max-gaps", 2147483647)), contig_start_mode=int(flags.get("--contig-start-mode", 2)),
               contig_end_mode=int(flags.get("--contig-end-mode", 2)), orf_start_mode=int(flags.get("--orf-start-mode", 1)),
               forward_frames=frames(flags.get("--forward-frames", "1,2,3")), reverse_frames=frames(flags.get("--reverse-frames", "1,2,3")),
               translation_table=int(flags.get("--translation-table", 1)), use_all_table_starts=int(flags.get("--use-all-table-starts", 0)))


def extractorfs(db, op, translate):
    """Returns (fragment DB keyed 0..n-1, orf_info uint32 (n, 4): read key, fromPos, toPos, flags)."""
    from plass_b200.mmseqsdb import DB
    od, oo, ol, ok, oi = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    on, ob = C.c_uint64(), C.c_uint64()
    s = seqdb_struct(db)
    rc = lib().or_extractorfs(C.byref(s), C.byref(op), C.c_int(1 if translate else 0), C.byref(od), C.byref(oo), C.byref(ol), C.byref(ok),
                              C.byref(on), C.byref(ob), C.byref(oi))
    assert rc == 0, rc
    n = on.value
    out = DB(_take(od, ob.value, np.dtype("u1")), _take(ok, n, np.dtype("<u4")), _take(oo, n, np.dtype("<u8")),
             _take(ol, n, np.dtype("<u4")), 0 if translate else 1)
    return out, _take(oi, 4 * n, np.dtype("<u4")).reshape(n, 4)


def translatenucs(db, flags=None, max_seq_len=65535):
    from plass_b200.mmseqsdb import DB
    od, oo, ol, ok = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    on, ob = C.c_uint64(), C.c_uint64()
    s = seqdb_struct(db)
    fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    rc = lib().or_translatenucs(C.byref(s), None if fl is None else C.c_void_p(fl.ctypes.data), C.c_int(max_seq_len),
                                C.byref(od), C.byref(oo), C.byref(ol), C.byref(ok), C.byref(on), C.byref(ob))
    assert rc == 0, rc
    n = on.value
    return DB(_take(od, ob.value, np.dtype("u1")), _take(ok, n, np.dtype("<u4")), _take(oo, n, np.dtype("<u8")), _take(ol, n, np.dtype("<u4")), 0)


def orf_flags_from_headers(hdr_db, keys):
    """translatenucs.cpp:58-63: addStopAtStart = !incompleteStart, addStopAtEnd = !incompleteEnd from the ORF header of each key."""
    h = hdr_db.entries_by_key()
    out = np.zeros(len(keys), dtype=np.uint8)
    for i, k in enumerate(keys):
        cols = h[int(k)].decode().split()
        complete = int(cols[2]) if len(cols) >= 3 else 0
        out[i] = (0 if (complete & 1) else 1) | (0 if (complete & 2) else 2)
    return out


def orf_header_entries(info):
    """Orf::writeOrfHeader (mm/commons/Orf.cpp:445-462): {new key: b"readKey\tfrom+len[\tflags]\n"}."""
    out = {}
    for k, (key, f, t, fl) in enumerate(info.tolist()):
        e = "%d\t%d%s%d" % (key, f, "+" if f < t else "-", abs(f - t))
        if fl:
            e += "\t%d" % fl
        out[k] = (e + "\n").encode()
    return out


def cyclecheck(db, max_seq_len, kmer_size=22):
    """cyclecheck: splitDiagonal per sequence (0 = not circular)."""
    split = np.zeros(db.n, dtype=np.uint32)
    s = seqdb_struct(db)
    rc = lib().or_cyclecheck(C.byref(s), C.c_int(max_seq_len), C.c_int(kmer_size), C.c_void_p(split.ctypes.data))
    assert rc == 0, rc
    return split


def cycle_db(db, split, chop):
    """The DB cyclecheck writes for the given split diagonals (cyclecheck.cpp:249-259)."""
    from plass_b200.mmseqsdb import DB
    idx = np.nonzero(split)[0]
    parts, lens = [], []
    for i in idx:
        o, l = int(db.offsets[i]), int(db.lens[i])
        e = (bytes(db.data[o:o + int(split[i])]) + b"\n\0") if chop else bytes(db.data[o:o + l])
        parts.append(e); lens.append(len(e))
    lens = np.array(lens, dtype=np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64) if len(lens) else np.zeros(0, dtype=np.uint64)
    return DB(np.frombuffer(b"".join(parts), dtype=np.uint8), db.keys[idx].copy(), offs, lens, db.dbtype)


def format_hits_by_rep(db_keys, hits):
    """Logical content of a prefilter DB: {key: entry bytes} (self line + hit lines)."""
    buf = C.create_string_buffer(128)
    out = {int(k): bytearray(b"%d\t0\t0\n" % int(k)) for k in db_keys}
    L = lib()
    for h in hits:
        n = L.or_format_hit(buf, C.c_uint32(int(h["target"])), C.c_int32(int(h["score"])), C.c_int32(int(h["diag"])))
        out[int(h["rep"])] += buf.raw[:n]
    return {k: bytes(v) for k, v in out.items()}


def format_alns_by_query(db_keys, alns):
    buf = C.create_string_buffer(256)
    out = {int(k): bytearray() for k in db_keys}
    L = lib()
    alns = np.ascontiguousarray(alns)
    base = alns.ctypes.data
    for i in range(len(alns)):
        n = L.or_format_aln(buf, C.c_void_p(base + i * ALN.itemsize))
        out[int(alns[i]["query"])] += buf.raw[:n]
    return {k: bytes(v) for k, v in out.items()}
