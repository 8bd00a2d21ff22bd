"""Diagnostic: the fused iteration twice on the same DB in one context, and once more with the later stages' scratch kept
out of the record buffers; reports whether hits, alignments and the new DB are identical.  python scripts/determinism_check.py [reads]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plass_b200 import api, synth, sharded  # noqa: E402
import bench  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000000
ctx = api.Context(0)
lib = api.load_library()
kp, rp, ep = api.default_km_params(False), api.default_rs_params(False), api.default_ex_params(False)
ddb = bench.fragments_gpu(ctx, synth.make_reads_fast(n_reads, seed=1))


def run(tag):
    out, hits, alns = ctx.assemble_iteration(ddb, kp, rp, ep, want_intermediates=True)
    h = out.download()
    out.free()
    res = {"tag": tag, "hits": len(hits), "alns": len(alns), "hit_ck": int(sharded.record_checksum(hits)), "aln_ck": int(sharded.record_checksum(alns)),
           "n": int(h.n), "bytes": int(h.data.nbytes), "lens_sum": int(np.asarray(h.lens, dtype=np.uint64).sum())}
    return res, h


r1, h1 = run("first")
r2, h2 = run("second")
lib.pg_debug_no_scratch_alias(ctx.handle, 1)
r3, h3 = run("no-alias")
r4, h4 = run("no-alias again")
out = {"reads": n_reads, "runs": [r1, r2, r3, r4]}
for name, a, b in (("first_vs_second", h1, h2), ("first_vs_noalias", h1, h3), ("noalias_vs_noalias", h3, h4)):
    same_len = bool(np.array_equal(a.lens, b.lens))
    out[name] = {"lens_equal": same_len, "entries_with_other_length": int((np.asarray(a.lens) != np.asarray(b.lens)).sum()) if a.n == b.n else -1,
                 "data_equal": bool(same_len and np.array_equal(a.data, b.data))}
print(json.dumps(out))
ctx.close()
