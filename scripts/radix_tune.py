#!/usr/bin/env python3
"""Micro-benchmark of the radix scatter pass on one B200: ms per pass and physical GB/s for the tile sizes and digit widths."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plass_b200 import api

ctx = api.Context(0)
lib = api.load_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 284000000
for bits in (8, 9, 10):
    for items in (16, 12) + ((8,) if bits == 8 else ()):
        ms = C.c_float()
        rc = lib.pg_debug_radix_bench_w(ctx.handle, C.c_uint64(n), C.c_int(items), C.c_int(4 if bits == 8 else 3), C.c_int(bits), C.byref(ms))
        assert rc == 0, lib.pg_last_error()
        print("bits=%2d items=%2d  %.3f ms/pass  %.0f GB/s physical (read+write)" % (bits, items, ms.value, n * 32 / 1e9 / (ms.value / 1e3)), flush=True)
ctx.close()
