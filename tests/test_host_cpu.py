"""CPU-only tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/plassgpu.h
declares, fails loudly without a GPU, the MMseqs DB reader/writer round-trips, the synthetic generators are
deterministic, and the sharding plan used for N > 1 GPUs is consistent (world_size-2 gloo run)."""
import ctypes
import os
import re
import subprocess
import sys
import numpy as np
import pytest

from common import ROOT, golden_case
from plass_b200 import mmseqsdb, synth


def test_library_exports_every_declared_symbol():
    from plass_b200 import api, build
    build.build()
    lib = api.load_library()
    header = open(os.path.join(ROOT, "include", "plassgpu.h")).read()
    declared = set(re.findall(r"\b(pg_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "libplassgpu.so does not export %s" % name
    assert set(api.EXPORTS) <= declared


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """sizeof / offsetof of every struct of include/plassgpu.h, as gcc lays it out, equals the ctypes mirrors in
    plass_b200/api.py and the numpy record types the results are read through."""
    from plass_b200 import api
    structs = {
        "pg_seqdb_view": (api.SeqDBView, None), "pg_km_params": (api.KmParams, None), "pg_rs_params": (api.RsParams, None),
        "pg_ex_params": (api.ExParams, None), "pg_timings": (api.Timings, None), "pg_orf_params": (api.OrfParams, None),
        "pg_hit": (None, api.HIT), "pg_aln": (None, api.ALN),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "plassgpu.h"', 'int main(void) {']
    for name, (ct, npd) in structs.items():
        fields = [f for f, _ in ct._fields_] if ct is not None else list(npd.names)
        lines.append('printf("%s %%zu", sizeof(%s));' % (name, name))
        for f in fields:
            lines.append('printf(" %s=%%zu", offsetof(%s, %s));' % (f, name, f))
        lines.append('printf("\\n");')
    lines += ['return 0; }']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "abi")
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), "-o", exe, str(src)], check=True)
    out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout.splitlines()
    assert len(out) == len(structs)
    for line in out:
        w = line.split()
        name, size = w[0], int(w[1])
        ct, npd = structs[name]
        offs = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in w[2:]}
        if ct is not None:
            assert ctypes.sizeof(ct) == size, name
            for f, _ in ct._fields_:
                assert getattr(ct, f).offset == offs[f], (name, f)
        else:
            assert npd.itemsize == size, name
            for f in npd.names:
                assert npd.fields[f][1] == offs[f], (name, f)


def test_no_cpu_fallback():
    from plass_b200 import api
    lib = api.load_library()
    if lib.pg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.PlassGpuError) as e:
        api.Context(0)
    assert "no CUDA device" in str(e.value)


def test_cli_errors_are_clean_failures(golden_root, tmp_path):
    """The reference ends a command with EXIT_FAILURE and a message (EXIT(), Application.cpp); so does the drop-in: an unknown flag
    or a flag value the GPU path does not implement is a hard error with exit status 1 -- not a fallback, and not an abort of
    the process although the CUDA start-up is already running in its helper thread -- and so is a machine without a GPU."""
    import subprocess
    cli = os.path.join(ROOT, "plass_b200", "plass_b200_cli")
    from common import golden_case
    d, man = golden_case("synth_aa", golden_root)
    seq = os.path.join(d, [x for x in man["steps"] if x["cmd"] == "kmermatcher"][0]["dbs"][0])
    out = str(tmp_path / "x")
    for args, needle in ((["kmermatcher", seq, out, "--bogus", "1"], "unknown parameter"),
                         (["kmermatcher", seq, out, "--spaced-kmer-mode", "1"], "not supported"),
                         (["rescorediagonal", seq, seq, out, out, "--rescore-mode", "1"], "not supported"),
                         (["kmermatcher", str(tmp_path / "missing"), out], "")):
        r = subprocess.run([cli] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 1 and "Error:" in r.stdout and needle in r.stdout, (args, r.returncode, r.stdout[-500:])
    from plass_b200 import api
    if api.load_library().pg_device_count() == 0:
        r = subprocess.run([cli, "kmermatcher", seq, out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stdout, (r.returncode, r.stdout[-500:])


def test_product_never_touches_the_oracle():
    pat = re.compile(r"oracle", re.I)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "plass_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for m in pat.finditer(txt):
                    line = txt[txt.rfind("\n", 0, m.start()) + 1: txt.find("\n", m.end())]
                    assert line.lstrip().startswith(("//", "#", "*", '"""', "written")) or "GENERATED by oracle" in line or "oracle/" in line and "//" in line, \
                        "%s references the oracle outside a comment: %s" % (f, line)


def test_mmseqsdb_roundtrip(tmp_path):
    seqs = [b"MKV", b"", b"ACDEFGHIK*"]
    p = str(tmp_path / "db")
    mmseqsdb.write_db(p, [5, 2, 9], seqs, 0)
    db = mmseqsdb.read_db(p)
    assert list(db.keys) == [2, 5, 9]
    assert db.entries_by_key() == {5: b"MKV", 2: b"", 9: b"ACDEFGHIK*"}
    mem = mmseqsdb.from_sequences([b"MKV", b"AC"], 0)
    assert mem.entry(0) == b"MKV\n" and list(mem.lens) == [5, 4] and list(mem.offsets) == [0, 5]


def test_golden_dbs_are_canonical(golden_root):
    d, man = golden_case("example_aa", golden_root)
    db = mmseqsdb.read_db(os.path.join(d, man["steps"][0]["dbs"][0]))
    assert db.n == 7255 and db.dbtype == 0
    assert np.all(np.diff(db.keys.astype(np.int64)) > 0)


def test_synth_is_seeded():
    a, b = synth.make_reads(500, seed=3), synth.make_reads(500, seed=3)
    assert np.array_equal(a, b) and not np.array_equal(a, synth.make_reads(500, seed=4))
    db = synth.protein_fragments(a)
    assert db.n > 300 and db.lens.min() >= 47 and bytes(db.data[int(db.offsets[1]) - 2:int(db.offsets[1])]) == b"\n\0"
    f = synth.make_reads_fast(500, seed=3)
    assert f.shape == (500, 150) and set(np.unique(f)) <= set(b"ACGT")


def test_sharding_plan_two_ranks_gloo(tmp_path):
    """world_size 2 over gloo: both ranks derive the same hash ranges / owner map and the all-to-all split
    sizes they exchange are consistent (plass_b200.sharded host logic, no GPU involved)."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", script], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("gloo-ok") == 2, r.stdout[-3000:]


def test_text_layer_round_trip_and_dbdiff(golden_root, tmp_path):
    """plass_b200_cli iotest: prefilter / alignment DBs of the reference -> records -> text must reproduce the reference's
    bytes (parallel index parse, parallel entry parse, parallel formatting + pwrite), for 1 and 4 host threads; dbdiff
    reports identical DBs as identical and a changed score / E-value as such.  No GPU involved."""
    import json
    import shutil
    import subprocess
    cli = os.path.join(ROOT, "plass_b200", "plass_b200_cli")
    from common import golden_case
    for case, names in (("synth_aa", ("pref_0", "aln_0")), ("synth_nt", ("pref_1", "aln_1"))):
        d, _ = golden_case(case, golden_root)
        for name in names:
            kind = "pref" if name.startswith("pref") else "aln"
            # result DBs are written with one data file per host thread (X.0 .. X.k, the layout the reference's own writer leaves
            # without a merge); a single non-empty file is named X.  The same output path is reused with fewer threads: no data
            # file of the earlier layout may survive.
            out = str(tmp_path / ("%s_%s" % (case, name)))
            for threads in ("4", "64", "1"):
                r = subprocess.run([cli, "iotest", kind, os.path.join(d, name), out, "--threads", threads], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                assert r.returncode == 0, r.stdout
                r = subprocess.run([cli, "dbdiff", out, os.path.join(d, name)], stdout=subprocess.PIPE, text=True)
                rep = json.loads(r.stdout.splitlines()[-1])
                assert r.returncode == 0 and rep["mismatching"] == 0 and rep["identical"] == rep["entries_b"] > 0, (case, name, threads, rep)
                parts = sorted(f for f in os.listdir(str(tmp_path)) if f.startswith(os.path.basename(out) + ".") and f.rsplit(".", 1)[1].isdigit())
                if threads == "1":
                    assert os.path.exists(out) and not parts, (threads, parts)
                else:
                    assert parts == ["%s.%d" % (os.path.basename(out), i) for i in range(len(parts))] and (len(parts) >= 2) != os.path.exists(out), parts
                    assert all(os.path.getsize(os.path.join(str(tmp_path), f)) > 0 for f in parts)
                assert mmseqsdb.read_db(out).entries_by_key() == mmseqsdb.read_db(os.path.join(d, name)).entries_by_key()
    # an empty result DB stays an empty data file + empty index (no per-thread files left behind)
    empty = str(tmp_path / "empty_pref")
    for ext in ("", ".index"):
        open(empty + ext, "wb").close()
    np.array([7], dtype="<i4").tofile(empty + ".dbtype")
    out = str(tmp_path / "empty_rt")
    r = subprocess.run([cli, "iotest", "pref", empty, out, "--threads", "4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert os.path.getsize(out) == 0 and os.path.getsize(out + ".index") == 0 and not os.path.exists(out + ".0")
    # sequence DBs through the writer of the GPU commands' results: back-to-back entries take the one-writer fast path (the buffer
    # is the data file), anything else goes entry by entry; both must reproduce the input DB
    d, man = golden_case("synth_aa", golden_root)
    step = [x for x in man["steps"] if x["cmd"] == "assembleresults"][0]
    for name in (step["dbs"][0], step["dbs"][2]):
        src = mmseqsdb.read_db(os.path.join(d, name))
        packed = str(tmp_path / ("packed_" + name))
        ent = src.entries_by_key()
        mmseqsdb.write_db(packed, sorted(ent), [ent[k] for k in sorted(ent)], src.dbtype)   # key order, back to back
        for inp in (os.path.join(d, name), packed):
            for threads in ("1", "4"):
                out = str(tmp_path / ("seq_rt_" + name))
                r = subprocess.run([cli, "iotest", "seq", inp, out, "--threads", threads], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                assert r.returncode == 0, r.stdout
                got = mmseqsdb.read_db(out)
                assert got.dbtype == src.dbtype and got.entries_by_key() == src.entries_by_key(), (name, inp, threads)
    # every sequence identity the reference can print ("0.000" .. "0.999", "1.00"): iotest cross-checks the column parser's
    # (float) (digits / 1000) against strtof on every line it reads and fails on a difference
    lines = ["%d\t%d\t%s\t1.000E-05\t0\t49\t50\t0\t49\t50\n" % (k, 50 + k % 7, "0.%03d" % k if k < 1000 else "1.00") for k in range(1001)]
    ident = str(tmp_path / "ident_aln")
    mmseqsdb.write_db(ident, [0, 1], ["".join(lines[:500]).encode(), "".join(lines[500:]).encode()], 5)
    out = str(tmp_path / "ident_rt")
    r = subprocess.run([cli, "iotest", "aln", ident, out, "--threads", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert len(mmseqsdb.read_db(out).entry(0).splitlines()) == 500
    # a flipped sign and a changed last E-value digit are found and classified
    d, _ = golden_case("synth_nt", golden_root)
    for name, mode, old, new in (("pref_0", "pref", b"\t-", b"\t"), ("aln_0", "aln", b"E-", b"E-")):
        src = os.path.join(d, name)
        dst = str(tmp_path / ("mut_" + name))
        for ext in ("", ".index", ".dbtype"):
            shutil.copy(src + ext, dst + ext)
        data = bytearray(open(dst, "rb").read())
        if mode == "pref":
            i = data.index(b"\t-")
            data[i + 1:i + 2] = b"+"                      # "-12" -> "+12": same magnitude, other strand
        else:
            i = data.index(b"E-") - 1
            data[i] = data[i] + 1 if data[i] < ord("9") else data[i] - 1
        open(dst, "wb").write(bytes(data))
        exact = json.loads(subprocess.run([cli, "dbdiff", dst, src], stdout=subprocess.PIPE, text=True).stdout.splitlines()[-1])
        tol = json.loads(subprocess.run([cli, "dbdiff", dst, src, "--mode", mode], stdout=subprocess.PIPE, text=True).stdout.splitlines()[-1])
        assert exact["mismatching"] == 1 and tol["mismatching"] == 0 and tol["tolerated"] == 1, (name, exact, tol)


def test_balanced_bounds_cpp_equals_python():
    """The equal-work cut of the representative key space: the C++ host arithmetic of pg_shard_iteration
    (pg_shard_balanced_bounds) against the Python mirror used by the emulated-rank tests."""
    import ctypes as C
    from plass_b200 import api, sharded
    lib = api.load_library()
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        for max_key in (10, 4095, 1000003, 83790975):
            hist = (rng.random(4096) * 1000 * np.exp(-np.arange(4096) / 300.0)).astype(np.uint64)
            out = (C.c_uint32 * (world + 1))()
            assert lib.pg_shard_balanced_bounds(C.c_void_p(hist.ctypes.data), 4096, C.c_uint32(max_key), world, C.c_double(4.0), out) == 0
            assert [int(x) for x in out] == [int(x) for x in sharded.balanced_bounds(hist, max_key, world)], (world, max_key)
