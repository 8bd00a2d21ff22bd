"""Micro-benchmark of the 256-bin radix pass variants on one B200 (device-resident pseudo-random 16-byte records):
ms per scatter pass and the DRAM rate it implies (one read + one write of every record).
  python scripts/radix_modes.py [records]      # default 284 M (the 5 M-read workload), 3 passes"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plass_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 284000000
ctx = api.Context(0)
lib = api.load_library()
out = {}
for mode in (0, 1, 2, 3):
    lib.pg_debug_set_radix_mode(mode)
    best = None
    for _ in range(3):
        ms = C.c_float()
        rc = lib.pg_debug_radix_bench(ctx.handle, C.c_uint64(n), 12, 3, C.byref(ms))
        assert rc == 0, lib.pg_last_error()
        best = ms.value if best is None else min(best, ms.value)
    out["mode%d" % mode] = {"ms_per_pass": best, "GBps": 32.0 * n / 1e9 / (best / 1e3)}
lib.pg_debug_set_radix_mode(1)
print(json.dumps({"records": n, **out}))
ctx.close()
