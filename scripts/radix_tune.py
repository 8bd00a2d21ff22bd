#!/usr/bin/env python3
"""Micro-benchmark of the radix scatter pass on one B200: ms per pass and physical GB/s for the three tile sizes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plass_b200 import api

ctx = api.Context(0)
lib = api.load_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 284000000
for items in (16, 12, 8):
    ms = C.c_float()
    rc = lib.pg_debug_radix_bench(ctx.handle, C.c_uint64(n), C.c_int(items), C.c_int(4), C.byref(ms))
    assert rc == 0, lib.pg_last_error()
    print("items=%2d  %.3f ms/pass  %.0f GB/s physical (read+write)" % (items, ms.value, n * 32 / 1e9 / (ms.value / 1e3)), flush=True)
ctx.close()
