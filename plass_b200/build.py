"""Builds libplassgpu.so (hand-written CUDA for sm_100a + the C ABI) and the host CLI, in-tree.

  python -m plass_b200.build          # or: from plass_b200.build import build; build()

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libplassgpu.so")
CLI = os.path.join(HERE, "plass_b200_cli")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v"]

# file -> extra flags.  The E-value / identity arithmetic must follow the CPU's operation order:
# no fused multiply-add contraction in those translation units.
CUDA_SOURCES = {
    "pg_api.cu": [],
    "pg_radix.cu": [],
    "pg_scan.cu": [],
    "pg_kmermatch.cu": ["-fmad=false"],   # (float) budget kmersPerSeq - 1 + scale * L: product and sum rounded separately (kmermatcher.cpp:223)
    "pg_rescore.cu": ["-fmad=false"],
    "pg_extend.cu": ["-fmad=false"],
    "pg_next.cu": ["-fmad=false"],
    "pg_orf.cu": [],
    "pg_shard.cu": [],
}
HOST_SOURCES = ["host/cli.cpp", "host/mmdb.cpp", "host/commands.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "plassgpu.h"))
    hd = os.path.join(CSRC, "host")
    if os.path.isdir(hd):
        hs += [os.path.join(hd, f) for f in os.listdir(hd) if f.endswith(".h")]
    return hs


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = _headers()
    jobs = []
    objs = []
    for src, extra in CUDA_SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if _newer(o, [s] + headers):
            jobs.append(([NVCC] + ARCH + COMMON + extra + ["-c", s, "-o", o], o + ".log"))
    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(lambda j: _run(*j), jobs))
    if verbose:
        for o in outs:
            print(o)
    if _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"], os.path.join(OBJ, "link.log"))
    host = [os.path.join(CSRC, h) for h in HOST_SOURCES]
    if all(os.path.exists(h) for h in host) and _newer(CLI, host + headers + [LIB]):
        _run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wall", "-I" + os.path.join(os.path.dirname(HERE), "include"), "-o", CLI] + host +
             ["-L" + HERE, "-lplassgpu", "-Wl,-rpath,$ORIGIN", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"],
             os.path.join(OBJ, "cli.log"))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
