"""Drop-in tests (-m gpu): the host CLI (plass_b200_cli) reads / writes the MMseqs2 on-disk format and
(a) reproduces every golden DB of the reference step by step, (b) slots into the reference's own
`plass assemble` / `penguin nuclassemble` shell workflow through scripts/plass_gpu and yields the
same intermediate DBs and the same final FASTA as the unmodified reference run on the same box."""
import os
import subprocess
import numpy as np
import pytest

from common import golden_case, ROOT
from plass_b200 import mmseqsdb, synth
from test_oracle_vs_reference import assert_same_entries

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "plass_b200", "plass_b200_cli")
REF = os.path.join(ROOT, "oracle", "_ref", "bin")


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    assert r.returncode == 0, "%s\n%s" % (" ".join(cmd), r.stdout[-3000:])
    return r.stdout


def aln_entries_close(got, want):
    """alignment entries: all columns identical except that the E-value column may differ in the last printed digit."""
    assert set(got) == set(want)
    n_diff = 0
    for k in want:
        if got[k] == want[k]:
            continue
        gl, wl = got[k].decode().splitlines(), want[k].decode().splitlines()
        assert len(gl) == len(wl), k
        for a, b in zip(gl, wl):
            ca, cb = a.split("\t"), b.split("\t")
            assert ca[:3] == cb[:3] and ca[4:] == cb[4:], (k, a, b)
            assert abs(float(ca[3]) - float(cb[3])) <= 2e-3 * abs(float(cb[3])), (k, a, b)
            n_diff += 1
    return n_diff


@pytest.mark.parametrize("case", ["example_aa", "synth_nt"])
def test_cli_reproduces_golden_steps(case, golden_root, tmp_path):
    d, man = golden_case(case, golden_root)
    for s in man["steps"]:
        npos = {"kmermatcher": 2, "rescorediagonal": 4}.get(s["cmd"], 3)
        ins = [os.path.join(d, x) for x in s["dbs"][:npos - 1]]
        out = str(tmp_path / (s["dbs"][-1] + "_gpu"))
        run([CLI, s["cmd"]] + ins + [out] + s["args"])
        got, want = mmseqsdb.read_db(out), mmseqsdb.read_db(os.path.join(d, s["dbs"][-1]))
        assert got.dbtype == want.dbtype
        if s["cmd"] == "rescorediagonal":
            aln_entries_close(got.entries_by_key(), want.entries_by_key())
        else:
            assert_same_entries(got.entries_by_key(), want.entries_by_key(), "%s/%s via CLI" % (case, s["dbs"][-1]))


@pytest.mark.parametrize("tool,wf,iters", [("plass", "assemble", 2), ("penguin", "nuclassemble", 2)])
def test_workflow_dropin_matches_reference(tool, wf, iters, tmp_path):
    ref_bin = os.path.join(REF, tool)
    if not os.path.exists(ref_bin):
        pytest.skip("reference binary oracle/_ref/bin/%s not available on this box" % tool)
    fa = str(tmp_path / "reads.fasta")
    synth.write_fasta(fa, synth.make_reads(3000, seed=21))
    # --threads 1: the reference's upstream extractorfs hands out new keys with an atomic counter, so with several
    # threads two runs number the fragments differently; the hot-path steps themselves are thread-independent.
    common = ["--num-iterations", str(iters), "--remove-tmp-files", "0", "--delete-tmp-inc", "0", "--threads", "1"]
    run([ref_bin, wf, fa, str(tmp_path / "ref.fas"), str(tmp_path / "tmp_ref")] + common)
    env = dict(os.environ, PLASS_REF_BIN=ref_bin)
    log = run([os.path.join(ROOT, "scripts", "plass_gpu"), wf, fa, str(tmp_path / "gpu.fas"), str(tmp_path / "tmp_gpu")] + common, env=env)
    assert "plass_b200" in log or "Time for processing" in log
    tr, tg = str(tmp_path / "tmp_ref" / "latest"), str(tmp_path / "tmp_gpu" / "latest")
    for i in range(iters):
        for name in ("pref_%d" % i, "assembly_%d" % i):
            assert_same_entries(mmseqsdb.read_db(os.path.join(tg, name)).entries_by_key(),
                                mmseqsdb.read_db(os.path.join(tr, name)).entries_by_key(), "%s %s" % (wf, name))
        aln_entries_close(mmseqsdb.read_db(os.path.join(tg, "aln_%d" % i)).entries_by_key(),
                          mmseqsdb.read_db(os.path.join(tr, "aln_%d" % i)).entries_by_key())
    assert open(str(tmp_path / "gpu.fas"), "rb").read() == open(str(tmp_path / "ref.fas"), "rb").read()
