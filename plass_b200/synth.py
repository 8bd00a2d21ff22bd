"""Seeded synthetic inputs: a coding "genome" of uniformly random sense codons with a TAA
stop every 300 codons; reads = uniform start, fixed length, uniform substitutions, 50 % reverse-
complemented.  `protein_fragments` turns reads into the amino-acid fragment DB the assemble iteration
works on (the shape `extractorfs` + `translatenucs` produce upstream of the hot path: every stop-free
stretch of >= 45 codons (extractorfs --min-length 45, Assembler.cpp:21) in each of the six frames; that upstream step itself is SURVEY §8f "next #2"
and is not reproduced byte for byte here)."""
import numpy as np

from . import mmseqsdb

_STOPS = {"TAA", "TAG", "TGA"}
_SENSE = np.array([[ord(a), ord(b), ord(c)] for a in "ACGT" for b in "ACGT" for c in "ACGT"
                   if a + b + c not in _STOPS], dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b

_AA = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"   # standard code, TCAG order
_CODE = np.zeros(256, dtype=np.uint8)
for _i, _ch in enumerate(b"TCAG"):
    _CODE[_ch] = _i
_AA_LUT = np.frombuffer(_AA.encode(), dtype=np.uint8)


def make_genome(n_nt, rng):
    n_codons = n_nt // 3 + 1
    g = _SENSE[rng.integers(0, len(_SENSE), n_codons)]
    g[299::300] = np.frombuffer(b"TAA", dtype=np.uint8)
    return g.reshape(-1)[:n_nt]


def make_reads(n_reads, read_len=150, coverage=20.0, sub_rate=0.005, seed=1):
    """(n_reads, read_len) uint8 array of ASCII nucleotides."""
    rng = np.random.default_rng(seed)
    n_nt = max(int(n_reads * read_len / coverage), read_len * 2)
    g = make_genome(n_nt, rng)
    starts = rng.integers(0, n_nt - read_len + 1, n_reads)
    reads = g[starts[:, None] + np.arange(read_len)[None, :]]
    sub = rng.random(reads.shape) < sub_rate
    if sub.any():
        reads[sub] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(sub.sum()))]
    rc = rng.random(n_reads) < 0.5
    reads[rc] = _COMP[reads[rc][:, ::-1]]
    return reads


def make_reads_fast(n_reads, read_len=150, coverage=20.0, sub_rate=0.005, seed=1):
    """Same model as make_reads, sized for millions of reads: substitutions are drawn as a sparse set of
    positions instead of one uniform per base (different random stream, same distribution)."""
    rng = np.random.default_rng(seed)
    n_nt = max(int(n_reads * read_len / coverage), read_len * 2)
    g = make_genome(n_nt, rng)
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    windows = np.lib.stride_tricks.sliding_window_view(g, read_len)
    for s in range(0, n_reads, 1000000):
        e = min(n_reads, s + 1000000)
        starts = rng.integers(0, n_nt - read_len + 1, e - s)
        blk = windows[starts]
        rc = rng.random(e - s) < 0.5
        blk[rc] = _COMP[blk[rc][:, ::-1]]
        reads[s:e] = blk
    n_sub = rng.binomial(n_reads * read_len, sub_rate)
    pos = rng.integers(0, n_reads * read_len, n_sub)
    reads.reshape(-1)[pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n_sub)]
    return reads


def write_fasta(path, reads):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i)
            f.write(r.tobytes())
            f.write(b"\n")


def _translate(codes):
    """codes: (n, 3k) array of 2-bit codes (TCAG order) -> (n, k) ASCII amino acids."""
    c = codes.reshape(codes.shape[0], -1, 3).astype(np.uint8)
    return _AA_LUT[(c[:, :, 0] << 4) | (c[:, :, 1] << 2) | c[:, :, 2]]


def _fragments_of_chunk(r, min_len):
    """Fragments of one block of reads: (data pieces, length pieces) in strand-major, frame-minor order."""
    datas, lens = [], []
    for strand in (r, _COMP[r[:, ::-1]]):
        codes = _CODE[strand]
        for f in range(3):
            k = (codes.shape[1] - f) // 3
            aa = _translate(codes[:, f:f + 3 * k])
            # runs between stops, per row: append a stop column so runs never cross rows
            stop = np.ones((aa.shape[0], k + 1), dtype=bool)
            stop[:, :k] = aa == ord("*")
            flat_stop = stop.reshape(-1)
            flat = np.concatenate([aa, np.full((aa.shape[0], 1), ord("*"), np.uint8)], axis=1).reshape(-1)
            ends = np.flatnonzero(flat_stop)
            starts = np.concatenate([[0], ends[:-1] + 1])
            ln = ends - starts
            sel = ln >= min_len
            st, ln = starts[sel], ln[sel]
            if len(st) == 0:
                continue
            tot = int(ln.sum()) + 2 * len(ln)
            out = np.empty(tot, dtype=np.uint8)
            o = np.zeros(len(ln), dtype=np.int64)
            o[1:] = np.cumsum(ln[:-1] + 2)
            # gather residues: index = start[i] + (pos - o[i]) for pos within the fragment
            idx = np.repeat(st - o, ln + 2) + np.arange(tot)
            body = np.repeat(np.arange(len(ln)), ln + 2)
            within = np.arange(tot) - o[body]
            is_nl = within == ln[body]
            is_nul = within == ln[body] + 1
            idx[is_nl | is_nul] = 0
            out[:] = flat[idx]
            out[is_nl] = 10
            out[is_nul] = 0
            datas.append(out)
            lens.append((ln + 2).astype(np.uint32))
    data = np.concatenate(datas) if datas else np.zeros(0, np.uint8)
    lens = np.concatenate(lens) if lens else np.zeros(0, np.uint32)
    return data, lens


_POOL_READS = None


def _pool_chunk(args):
    s, e, min_len = args
    return _fragments_of_chunk(_POOL_READS[s:e], min_len)


def protein_fragments(reads, min_len=45, chunk=500000, workers=1):
    """Six-frame translation + split at stops; returns an in-memory amino-acid sequence DB.
    workers > 1: the blocks of `chunk` reads are translated by forked worker processes (same result, same order)."""
    global _POOL_READS
    spans = [(s, min(len(reads), s + chunk), min_len) for s in range(0, len(reads), chunk)]
    if workers > 1 and len(spans) > 1:
        import multiprocessing as mp
        _POOL_READS = reads
        try:
            with mp.get_context("fork").Pool(min(workers, len(spans))) as pool:
                parts = pool.map(_pool_chunk, spans, chunksize=1)
        finally:
            _POOL_READS = None
    else:
        parts = [_fragments_of_chunk(reads[s:e], min_len) for s, e, _ in spans]
    datas = [p[0] for p in parts]
    lens = [p[1] for p in parts]
    data = np.concatenate(datas) if datas else np.zeros(0, np.uint8)
    lens = np.concatenate(lens) if lens else np.zeros(0, np.uint32)
    offsets = np.zeros(len(lens), dtype=np.uint64)
    if len(lens):
        offsets[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    return mmseqsdb.DB(data, np.arange(len(lens), dtype=np.uint32), offsets, lens, mmseqsdb.DBTYPE_AMINO_ACIDS)


def nucleotide_db(reads):
    n, L = reads.shape
    data = np.empty((n, L + 2), dtype=np.uint8)
    data[:, :L] = reads
    data[:, L] = 10
    data[:, L + 1] = 0
    lens = np.full(n, L + 2, dtype=np.uint32)
    offsets = np.arange(n, dtype=np.uint64) * np.uint64(L + 2)
    return mmseqsdb.DB(data.reshape(-1), np.arange(n, dtype=np.uint32), offsets, lens, mmseqsdb.DBTYPE_NUCLEOTIDES)
