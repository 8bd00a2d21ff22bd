"""Seeded synthetic read generator: coding "genome" of uniformly random sense codons
with a TAA stop every 300 codons; reads = uniform start, fixed length, uniform substitutions,
50 % reverse-complemented.  Used by make_golden.py (via FASTA) and by bench.py (in memory)."""
import numpy as np

_STOPS = {"TAA", "TAG", "TGA"}
_SENSE = np.array([[ord(a), ord(b), ord(c)] for a in "ACGT" for b in "ACGT" for c in "ACGT"
                   if a + b + c not in _STOPS], dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def make_genome(n_nt, rng):
    n_codons = n_nt // 3 + 1
    g = _SENSE[rng.integers(0, len(_SENSE), n_codons)]
    g[299::300] = np.frombuffer(b"TAA", dtype=np.uint8)
    return g.reshape(-1)[:n_nt]


def make_reads(n_reads, read_len=150, coverage=20.0, sub_rate=0.005, seed=1):
    """Returns a (n_reads, read_len) uint8 array of ASCII nucleotides."""
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


def write_fasta(path, reads):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i)
            f.write(r.tobytes())
            f.write(b"\n")
