#!/usr/bin/env python3
"""Times the GPU six-frame fragment pipeline (extractorfs x 2 + translatenucs fused + concatdbs) on synthetic reads."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plass_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000000
ctx = api.Context(0)
reads = synth.make_reads_fast(n, seed=1)
dn = ctx.upload(synth.nucleotide_db(reads))
for rep in range(2):
    parts = []
    lo, _ = ctx.extractorfs(dn, api.orf_params_long(), translate=True, want_info=False); parts.append(ctx.timings()["total_ms"])
    st, _ = ctx.extractorfs(dn, api.orf_params_start(), translate=True, want_info=False); parts.append(ctx.timings()["total_ms"])
    cat = ctx.concat(lo, st); parts.append(ctx.timings()["total_ms"])
    nf = cat.n
    lo.free(); st.free(); cat.free()
print("reads %d fragments %d: long %.2f ms, start %.2f ms, concat %.2f ms" % (n, nf, parts[0], parts[1], parts[2]))
ctx.close()
