"""Minimal reader/writer for the MMseqs2 on-disk DB triple (data, .index, .dbtype).

Host-side helper used by the tests, the golden-fixture generator and bench.py; the product's own
DB layer is the C++ one in plass_b200/csrc/host/ (this module mirrors its behaviour in Python).
Format (reference: lib/mmseqs/src/commons/DBReader.cpp:173-253,770-831, DBWriter.cpp:193-252,522-614):
  X or X.0..X.k   concatenated entry bytes; every entry ends with '\\0'
  X.index         text lines  key \\t offset \\t length   (length includes the trailing '\\0';
                  offsets run over the concatenation of the split data files; NOT guaranteed key-sorted)
  X.dbtype        4-byte little-endian int (0 aa, 1 nt, 5 alignment, 7 prefilter, 14 prefilter-rev)
"""
import os
import numpy as np

DBTYPE_AMINO_ACIDS = 0
DBTYPE_NUCLEOTIDES = 1
DBTYPE_ALIGNMENT_RES = 5
DBTYPE_PREFILTER_RES = 7
DBTYPE_PREFILTER_REV_RES = 14


class DB:
    """In-memory DB: `data` (uint8), `keys` (uint32), `offsets` (uint64), `lens` (uint32), sorted by key."""

    def __init__(self, data, keys, offsets, lens, dbtype):
        self.data = data
        self.keys = keys
        self.offsets = offsets
        self.lens = lens
        self.dbtype = dbtype

    @property
    def n(self):
        return len(self.keys)

    def entry(self, i):
        """Entry bytes without the trailing NUL."""
        o = int(self.offsets[i])
        return self.data[o:o + int(self.lens[i]) - 1].tobytes()

    def entries_by_key(self):
        return {int(k): self.entry(i) for i, k in enumerate(self.keys)}


def _data_files(path):
    if os.path.exists(path):
        return [path]
    files = []
    i = 0
    while os.path.exists("%s.%d" % (path, i)):
        files.append("%s.%d" % (path, i))
        i += 1
    if not files:
        raise FileNotFoundError(path)
    return files


def read_db(path):
    parts = [np.fromfile(f, dtype=np.uint8) for f in _data_files(path)]
    data = np.concatenate(parts) if len(parts) > 1 else parts[0]
    idx = np.loadtxt(path + ".index", dtype=np.uint64, ndmin=2) if os.path.getsize(path + ".index") else np.zeros((0, 3), np.uint64)
    order = np.argsort(idx[:, 0], kind="stable")
    idx = idx[order]
    dbtype = int(np.fromfile(path + ".dbtype", dtype="<i4")[0]) & 0x7FFFFFFF
    return DB(data, idx[:, 0].astype(np.uint32), idx[:, 1].astype(np.uint64), idx[:, 2].astype(np.uint32), dbtype & 0xFFFF)


def write_db(path, keys, entries, dbtype):
    """Write a canonical single-file DB: entries (bytes, without NUL) in key order."""
    order = np.argsort(np.asarray(keys, dtype=np.uint64), kind="stable")
    off = 0
    with open(path, "wb") as fd, open(path + ".index", "w") as fi:
        for j in order:
            e = entries[j]
            fd.write(e)
            fd.write(b"\0")
            fi.write("%d\t%d\t%d\n" % (int(keys[j]), off, len(e) + 1))
            off += len(e) + 1
    np.array([dbtype], dtype="<i4").tofile(path + ".dbtype")


def canonicalize(src, dst):
    """Rewrite DB `src` (possibly split, unsorted index) as canonical single-file DB `dst`."""
    db = read_db(src)
    write_db(dst, db.keys, [db.entry(i) for i in range(db.n)], db.dbtype)


def from_sequences(seqs, dbtype, keys=None):
    """Build an in-memory sequence DB from a list of bytes objects (residues only)."""
    n = len(seqs)
    lens = np.array([len(s) + 2 for s in seqs], dtype=np.uint32)
    offsets = np.zeros(n, dtype=np.uint64)
    if n:
        offsets[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    data = np.frombuffer(b"".join(s + b"\n\0" for s in seqs), dtype=np.uint8).copy()
    if keys is None:
        keys = np.arange(n, dtype=np.uint32)
    return DB(data, np.asarray(keys, dtype=np.uint32), offsets, lens, dbtype)
