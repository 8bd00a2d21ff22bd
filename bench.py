#!/usr/bin/env python3
"""bench.py -- reads/s through one full assemble iteration (kmermatcher + rescorediagonal + assembleresults).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--reads R]

Workload (BASELINE.json configs[2] / [3], the configuration the north-star metric is quoted on): R = 50 M synthetic
150 bp protein-coding reads (seed 1, 20x coverage, 0.5 % substitutions, 50 % reverse-complemented) -> the six-frame
fragment DB `aa_6f_start_long` exactly as data/assemble.sh:41-77 builds it (extractorfs x 2 + translatenucs x 2 +
concatdbs; built here by the repo's own six-frame GPU pipeline, pg_extractorfs) -> one `plass assemble` iteration with
the workflow defaults (k = 14, 13-letter alphabet, --min-seq-id 0.9, -e 1e-5).  A "step" is one pass of the hot path
over that fragment DB.  N > 1 (torchrun): STRONG scaling -- the same 50 M reads sharded over the N ranks.

  value         whole-job reads/s with the fragment DB already resident in HBM (device-resident fused iteration)
  e2e           the same iteration through the C ABI with HOST buffers: pinned H2D of the DB, D2H of the prefilter hits,
                the alignments and the new sequence DB inside the timed region
  roofline      dominant kernel = the radix scatter passes of sort #1; achieved = algorithmic bytes of the sort
                (one read + one write of every 16-byte record) / summed scatter time; traffic from the
                round's ncu capture (profiles/roofline_capture.json)
  cpu_baseline  the UNMODIFIED reference binary (oracle/_ref/bin/plass: kmermatcher, rescorediagonal, assembleresults
                sub-commands, all host threads) on a bounded sample of the same workload
  parity        the sample the reference just processed goes through the GPU drop-in commands (plass_b200_cli) and
                pref / aln / assembly are compared key -> entry bytes with the reference's files; any mismatch fails
                the run
  dropin        wall-clock of the drop-in commands (process start to exit, DB files in and out) next to
                the reference's on the same DB

--impl reference times only the reference arm (CPU) on the same config and prints it as the line's value.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from plass_b200 import mmseqsdb, synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "plass")
CLI = os.path.join(ROOT, "plass_b200", "plass_b200_cli")
METRIC = "reads/sec per assemble iteration (kmermatcher+rescorediagonal+assembleresults)"
KM_FLAGS = "--sub-mat nucl:nucleotide.out,aa:blosum62.out --alph-size 13 --min-seq-id 0.9 --kmer-per-seq 60 " \
           "--spaced-kmer-mode 0 --kmer-per-seq-scale nucl:0.200,aa:0.000 --adjust-kmer-len 0 --mask 0 --mask-lower-case 0 " \
           "--cov-mode 0 -k 14 -c 0 --max-seq-len 65535 --hash-shift 67 --split-memory-limit 0 --include-only-extendable 0 " \
           "--ignore-multi-kmer 1 --compressed 0 -v 3"
RS_FLAGS = "--sub-mat nucl:nucleotide.out,aa:blosum62.out --rescore-mode 3 --wrapped-scoring 0 --filter-hits 0 -e 1e-05 -c 0 -a 0 " \
           "--cov-mode 0 --min-seq-id 0.9 --min-aln-len 0 --seq-id-mode 0 --add-self-matches 0 --sort-results 0 --db-load-mode 0 " \
           "--compressed 0 -v 3"
EX_FLAGS = "--min-seq-id 0.9 --max-seq-len 65535 --keep-target 1 -v 3 --rescore-mode 3"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(reads):
    """The line's `config`: identical for both arms (it names the workload, not what a run found in it)."""
    cfg = {5000000: " (BASELINE.json configs[1])", 50000000: " (BASELINE.json configs[2] / configs[3], one iteration)"}.get(reads, "")
    return {"workload": "%gM synthetic 150bp coding reads -> six-frame fragment DB aa_6f_start_long (extractorfs x2 + translatenucs x2 + concatdbs), "
                        "one plass assemble iteration: k=14, alph 13, --min-seq-id 0.9, -e 1e-5%s" % (reads / 1e6, cfg),
            "reads": reads, "l2": "inputs_larger_than_l2 (GBs of k-mer records per step, 126 MB L2)"}


# ---- workload ---------------------------------------------------------------------------------------------------
def concat_dbs(a, b):
    """concatdbs without --preserve-keys (data/assemble.sh:68-77): a's entries, then b's, keys renumbered."""
    lens = np.concatenate([a.lens, b.lens])
    offsets = np.zeros(len(lens), dtype=np.uint64)
    if len(lens):
        offsets[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    return mmseqsdb.DB(np.concatenate([a.data, b.data]), np.arange(len(lens), dtype=np.uint32), offsets, lens, a.dbtype)


_ORACLE_READS = None


def _oracle_chunk(span):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    s, e = span
    db = synth.nucleotide_db(_ORACLE_READS[s:e])
    long_p = ob.orf_params_from_flags({"--min-length": "45", "--max-length": "32734", "--max-gaps": "0", "--contig-start-mode": "2",
                                       "--contig-end-mode": "2", "--orf-start-mode": "0"})
    start_p = ob.orf_params_from_flags({"--min-length": "20", "--max-length": "45", "--max-gaps": "0", "--contig-start-mode": "1",
                                        "--contig-end-mode": "0", "--orf-start-mode": "0"})
    lo, _ = ob.extractorfs(db, long_p, True)
    st, _ = ob.extractorfs(db, start_p, True)
    return (np.array(lo.data), np.array(lo.lens)), (np.array(st.data), np.array(st.lens))


def fragments_cpu(reads, workers):
    """aa_6f_start_long of `reads` WITHOUT a GPU (reference arm): the oracle's restatement of extractorfs + translatenucs
    (oracle/oracle_next.cpp, pinned against the reference binary on tests/golden/orf_aa), forked over the host cores.
    Parameters = EXTRACTORFS_LONG_PAR / EXTRACTORFS_START_PAR of src/workflow/Assembler.cpp:114-128."""
    global _ORACLE_READS
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    ob.build()
    chunk = 100000
    spans = [(s, min(len(reads), s + chunk)) for s in range(0, len(reads), chunk)]
    _ORACLE_READS = reads
    try:
        if workers > 1 and len(spans) > 1:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(min(workers, len(spans))) as pool:
                parts = pool.map(_oracle_chunk, spans, chunksize=1)
        else:
            parts = [_oracle_chunk(s) for s in spans]
    finally:
        _ORACLE_READS = None

    def join(idx):
        lens = np.concatenate([p[idx][1] for p in parts]).astype(np.uint32)
        data = np.concatenate([p[idx][0] for p in parts])
        offsets = np.zeros(len(lens), dtype=np.uint64)
        if len(lens):
            offsets[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        return mmseqsdb.DB(data, np.arange(len(lens), dtype=np.uint32), offsets, lens, 0)
    return concat_dbs(join(0), join(1))


def fragments_gpu(ctx, reads):
    """aa_6f_start_long of `reads` on the GPU (pg_extractorfs x 2 + pg_seqdb_concat); returns the device DB."""
    dn = ctx.upload(synth.nucleotide_db(reads))
    frag = ctx.six_frame_fragments(dn)
    dn.free()
    return frag


def write_db_fast(path, db):
    np.asarray(db.data).tofile(path)
    k = np.asarray(db.keys, dtype=np.uint64); o = np.asarray(db.offsets, dtype=np.uint64); l = np.asarray(db.lens, dtype=np.uint64)
    try:
        import pyarrow as pa
        import pyarrow.csv as pacsv
        pacsv.write_csv(pa.table({"k": k, "o": o, "l": l}), path + ".index", pacsv.WriteOptions(include_header=False, delimiter="\t"))
    except Exception:  # noqa: BLE001
        np.savetxt(path + ".index", np.stack([k, o, l], axis=1), fmt="%d", delimiter="\t")
    np.array([db.dbtype], dtype="<i4").tofile(path + ".dbtype")


# ---- reference binary ---------------------------------------------------------------------------------------------
def run_reference_iteration(seq, threads, workdir):
    """kmermatcher + rescorediagonal + assembleresults of the unmodified reference on DB `seq` (on disk); returns
    (seconds, per-step seconds); leaves pref / aln / asm in workdir."""
    for name in ("pref", "aln", "asm"):
        for f in os.listdir(workdir):
            if f == name or f.startswith(name + "."):
                os.unlink(os.path.join(workdir, f))
    env = dict(os.environ, MMSEQS_NUM_THREADS=str(threads))
    w = lambda x: os.path.join(workdir, x)  # noqa: E731
    cmds = [
        [REF_BIN, "kmermatcher", seq, w("pref")] + KM_FLAGS.split() + ["--threads", str(threads)],
        [REF_BIN, "rescorediagonal", seq, seq, w("pref"), w("aln")] + RS_FLAGS.split() + ["--threads", str(threads)],
        [REF_BIN, "assembleresults", seq, w("aln"), w("asm")] + EX_FLAGS.split() + ["--threads", str(threads)],
    ]
    per = []
    for c in cmds:
        t = time.time()
        r = subprocess.run(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
        if r.returncode != 0:
            raise RuntimeError("reference step failed: %s\n%s" % (" ".join(c[:2]), r.stdout.decode()[-2000:]))
        per.append(time.time() - t)
    return sum(per), per


def run_cli(args, threads):
    t = time.time()
    r = subprocess.run([CLI] + args + ["--threads", str(threads)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("plass_b200_cli %s failed:\n%s" % (args[0], r.stdout[-2000:]))
    return time.time() - t, r.stdout


def dbdiff(a, b, mode):
    r = subprocess.run([CLI, "dbdiff", a, b, "--mode", mode], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    line = [x for x in r.stdout.splitlines() if x.startswith("{")]
    if not line:
        raise RuntimeError("dbdiff %s %s failed:\n%s" % (a, b, r.stdout[-2000:]))
    return json.loads(line[-1])


def parity_and_dropin(workdir, seq, threads, ref_per):
    """The sample DB `seq` (on disk, already processed by the reference into workdir/pref|aln|asm) through the GPU drop-in
    commands; compares the three result DBs key -> entry bytes and reports the commands' wall-clock."""
    w = lambda x: os.path.join(workdir, x)  # noqa: E731
    flags = lambda s: s.split()  # noqa: E731
    # (a) the three commands as data/assemble.sh calls them, one process each
    t_km, o_km = run_cli(["kmermatcher", seq, w("g_pref")] + flags(KM_FLAGS), threads)
    t_rs, o_rs = run_cli(["rescorediagonal", seq, seq, w("g_pref"), w("g_aln")] + flags(RS_FLAGS), threads)
    t_ex, o_ex = run_cli(["assembleresults", seq, w("g_aln"), w("g_asm")] + flags(EX_FLAGS), threads)
    phases = {n: ([x for x in o.splitlines() if x.startswith("Phases:")] or [None])[-1] for n, o in (("kmermatcher", o_km), ("rescorediagonal", o_rs), ("assembleresults", o_ex))}
    d_pref, d_aln, d_asm = dbdiff(w("g_pref"), w("pref"), "exact"), dbdiff(w("g_aln"), w("aln"), "aln"), dbdiff(w("g_asm"), w("asm"), "exact")
    # (b) the same iteration fused in one process
    union = flags(KM_FLAGS) + ["--rescore-mode", "3", "--wrapped-scoring", "0", "--filter-hits", "0", "-e", "1e-05", "-a", "0", "--min-aln-len", "0",
                               "--seq-id-mode", "0", "--add-self-matches", "0", "--sort-results", "0", "--db-load-mode", "0", "--keep-target", "1"]
    t_fused, out_fused = run_cli(["assembleiteration", seq, w("f_pref"), w("f_aln"), w("f_asm")] + union, threads)
    f_pref, f_aln, f_asm = dbdiff(w("f_pref"), w("pref"), "exact"), dbdiff(w("f_aln"), w("aln"), "aln"), dbdiff(w("f_asm"), w("asm"), "exact")
    # (c) the wall clock of a process start varies from box to box and run to run (CUDA start-up: 0.1 - 2.5 s observed): every command is
    # timed twice, the faster run is reported and both are listed
    first = {"kmermatcher": t_km, "rescorediagonal": t_rs, "assembleresults": t_ex, "fused": t_fused}
    t_km2, o_km2 = run_cli(["kmermatcher", seq, w("g_pref")] + flags(KM_FLAGS), threads)
    t_rs2, o_rs2 = run_cli(["rescorediagonal", seq, seq, w("g_pref"), w("g_aln")] + flags(RS_FLAGS), threads)
    t_ex2, o_ex2 = run_cli(["assembleresults", seq, w("g_aln"), w("g_asm")] + flags(EX_FLAGS), threads)
    t_fused2, out_fused2 = run_cli(["assembleiteration", seq, w("f_pref"), w("f_aln"), w("f_asm")] + union, threads)
    second = {"kmermatcher": t_km2, "rescorediagonal": t_rs2, "assembleresults": t_ex2, "fused": t_fused2}
    pick = lambda a, oa, b, ob: (a, oa) if a <= b else (b, ob)  # noqa: E731
    (t_km, o_km), (t_rs, o_rs), (t_ex, o_ex) = pick(t_km, o_km, t_km2, o_km2), pick(t_rs, o_rs, t_rs2, o_rs2), pick(t_ex, o_ex, t_ex2, o_ex2)
    t_fused, out_fused = pick(t_fused, out_fused, t_fused2, out_fused2)
    phases = {n: ([x for x in o.splitlines() if x.startswith("Phases:")] or [None])[-1] for n, o in (("kmermatcher", o_km), ("rescorediagonal", o_rs), ("assembleresults", o_ex))}
    mism = sum(d["mismatching"] + d["only_in_a"] + d["only_in_b"] for d in (d_pref, d_aln, d_asm, f_pref, f_aln, f_asm))
    parity = {"sequences": d_pref["entries_b"], "mismatching_entries": int(mism),
              "evalue_last_digit_entries": int(d_aln["tolerated"]), "evalue_last_digit_lines": int(d_aln["tolerated_lines"]),
              "compared": "pref, aln, assembly DBs of plass_b200_cli (three commands and fused assembleiteration) vs oracle/_ref/bin/plass on the same "
                          "sample DB, key -> entry bytes (plass_b200_cli dbdiff); aln: E-value column may differ by one unit of its last printed digit",
              "entries": {"pref": d_pref["entries_b"], "aln": d_aln["entries_b"], "assembly": d_asm["entries_b"]}}
    ref_total = sum(ref_per)
    dropin = {"definition": "wall-clock of each command as a process (start to exit: CUDA init, DB open / index parse, upload, kernels, download, text "
                            "formatting, DB write), SURVEY.md 8d; every GPU command is run twice and the faster run counts (both in gpu_cli_runs_s; the "
                            "reference's times are those of its last run); tmp dir %s; %d host threads" % (workdir, threads),
              "reference_s": {"kmermatcher": ref_per[0], "rescorediagonal": ref_per[1], "assembleresults": ref_per[2], "iteration": ref_total},
              "gpu_cli_s": {"kmermatcher": t_km, "rescorediagonal": t_rs, "assembleresults": t_ex, "iteration": t_km + t_rs + t_ex},
              "gpu_cli_runs_s": [first, second],
              "gpu_cli_phases": phases,
              "gpu_cli_fused_s": t_fused, "gpu_cli_fused_phases": [x for x in out_fused.splitlines() if x.startswith("open + index parse")][-1:] or None,
              "speedup_three_commands": ref_total / (t_km + t_rs + t_ex), "speedup_fused": ref_total / t_fused}
    return parity, dropin


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args):
    """--impl reference: the reference's own CPU path on a bounded sample of the same workload, all host threads."""
    threads = os.cpu_count() or 1
    work = tempfile.mkdtemp(prefix="plass_b200_ref_")
    try:
        # calibration: the same generator at the same 20x coverage (a prefix of the big DB would have lower coverage,
        # fewer overlaps per read, and flatter the CPU), enlarged towards ~10 s of CPU work per step
        sample_reads = min(args.cpu_sample_reads, args.reads)
        seq = os.path.join(work, "seq")
        db = fragments_cpu(synth.make_reads_fast(sample_reads, seed=args.seed), threads)
        write_db_fast(seq, db)
        t, per = run_reference_iteration(seq, threads, work)
        log("[bench/reference] calibration: %d reads (%d fragments) in %.2f s" % (sample_reads, db.n, t))
        if t < 7.0 and sample_reads < args.reads:
            want = min(args.reads, max(sample_reads, (int(sample_reads * 10.0 / max(t, 0.05)) // 100000) * 100000))
            if want > sample_reads:
                sample_reads = want
                db = fragments_cpu(synth.make_reads_fast(sample_reads, seed=args.seed), threads)
                write_db_fast(seq, db)
        times = []
        for i in range(args.warmup + args.steps):
            t, per = run_reference_iteration(seq, threads, work)
            log("[bench/reference] step %d: %.2f s (kmermatcher %.2f, rescorediagonal %.2f, assembleresults %.2f)" % (i, t, per[0], per[1], per[2]))
            if i >= args.warmup:
                times.append(t)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    sec = float(np.mean(times))
    val = sample_reads / sec
    sample_desc = "%d reads (%d fragments of aa_6f_start_long) from the same generator at the same 20x coverage instead of %d reads; %d threads; tmp dir %s" % (
        sample_reads, db.n, args.reads, threads, tempfile.gettempdir())
    return {
        "metric": METRIC, "value": val, "unit": "reads/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1000.0,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32 (+f64 E-values)", "data": "synthetic",
        "config": workload_config(args.reads),
        "cpu_baseline": {"value": val, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": sample_desc},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


_REAL_STDOUT = None


def emit_json(line):
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner) also write to fd 1, so
    fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def pin(db):
    import torch
    return mmseqsdb.DB(torch.from_numpy(np.ascontiguousarray(db.data)).pin_memory().numpy(),
                       torch.from_numpy(np.ascontiguousarray(db.keys)).pin_memory().numpy(),
                       torch.from_numpy(np.ascontiguousarray(db.offsets).view(np.int64)).pin_memory().numpy().view(np.uint64),
                       torch.from_numpy(np.ascontiguousarray(db.lens).view(np.int32)).pin_memory().numpy().view(np.uint32), db.dbtype)


def roofline_traffic(nrec):
    """DRAM bytes of one scatter launch from the round's `ncu --set full` capture (profiles/roofline_capture.json:
    dram__bytes_read.sum + dram__bytes_write.sum and the records of the captured launch), scaled to this run's launch."""
    p = os.path.join(ROOT, "profiles", "roofline_capture.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    per_rec = (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["records"]
    return nrec * per_rec, "%s (%s): %.2f B per record and launch" % (d.get("kernel", "?"), d.get("source", p), per_rec)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=50000000, help="reads of the whole job (strong scaling over the ranks)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample-reads", type=int, default=1000000, help="reads in the bounded CPU-baseline / parity sample (enlarged towards ~10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference binary (also skips the parity gate and the drop-in timing)")
    ap.add_argument("--no-extras", action="store_true", help="skip the (untimed-for-the-metric) measurements of the neighbouring steps")
    ap.add_argument("--e2e-steps", type=int, default=0, help="pipelined end-to-end steps (default max(steps, 8))")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        emit_json(reference_arm(args))
        return 0

    import torch
    import torch.distributed as dist
    from plass_b200 import api
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    ctx = api.Context(local_rank)
    kp, rp, ep = api.default_km_params(False), api.default_rs_params(False), api.default_ex_params(False)
    threads = os.cpu_count() or 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- CPU baseline + parity gate + drop-in wall-clock on a bounded sample (rank 0, before the big buffers exist) ----
    cpu = parity = dropin = None
    if rank == 0 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        work = tempfile.mkdtemp(prefix="plass_b200_cpu_")
        try:
            seq = os.path.join(work, "seq")

            def make_sample(n_reads):
                d = fragments_gpu(ctx, synth.make_reads_fast(n_reads, seed=args.seed))
                h = d.download()
                d.free()
                write_db_fast(seq, h)
                return h.n
            sample_reads = min(args.cpu_sample_reads, args.reads)
            n_frag = make_sample(sample_reads)
            t, per = run_reference_iteration(seq, threads, work)
            log("[bench/cpu] calibration: %d reads (%d fragments) in %.2f s" % (sample_reads, n_frag, t))
            if t < 7.0 and sample_reads < args.reads:
                want = min(args.reads, max(sample_reads, (int(sample_reads * 10.0 / max(t, 0.05)) // 100000) * 100000))
                if want > sample_reads:
                    sample_reads = want
                    n_frag = make_sample(sample_reads)
                    t, per = run_reference_iteration(seq, threads, work)
            cpu = {"value": sample_reads / t, "unit": "reads/s", "cores": threads, "kind": "reference",
                   "sample": "%d reads (%d fragments of aa_6f_start_long), same generator and 20x coverage, one iteration, %.1f s (kmermatcher %.1f, rescorediagonal %.1f, "
                             "assembleresults %.1f)" % (sample_reads, n_frag, t, per[0], per[1], per[2])}
            log("[bench/cpu] reference: %s" % cpu["sample"])
            parity, dropin = parity_and_dropin(work, seq, threads, per)
            parity["reads"] = sample_reads
            log("[bench/parity] %s" % json.dumps(parity))
            log("[bench/dropin] %s" % json.dumps(dropin))
        finally:
            shutil.rmtree(work, ignore_errors=True)
        if parity["mismatching_entries"] != 0:
            raise SystemExit("PARITY FAILURE: %d entries of the GPU drop-in differ from the reference on the %d-read sample" % (parity["mismatching_entries"], sample_reads))

    # ---- the job's input: rank 0 builds it on its GPU, the other ranks receive it over NVLink ------------------------
    t_gen = time.time()
    runner = None
    if world > 1:
        from plass_b200 import sharded
        runner = sharded.ShardedIteration(ctx, dist, rank, world)
        ddb = runner.build_and_broadcast(lambda: fragments_gpu(ctx, synth.make_reads_fast(args.reads, seed=args.seed)))
    else:
        ddb = fragments_gpu(ctx, synth.make_reads_fast(args.reads, seed=args.seed))
    n_frag_total = ddb.n
    if rank == 0:
        log("[bench] %d reads -> %d fragments in HBM in %.1f s" % (args.reads, n_frag_total, time.time() - t_gen))

    # ---- device-resident arm (value) -------------------------------------------------------------
    peak, peak_src = load_peaks()
    tim = []
    # the clock sampler (nvidia-smi, one line every 50 ms) is started before the warm-up steps: the tool needs ~0.1 s
    # before its first line.  Warm-up and timed steps run the same kernels, so every sample is taken under the load
    # that is being measured.
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        out = runner.step(ddb, kp, rp, ep) if runner else ctx.assemble_iteration(ddb, kp, rp, ep)[0]
        out.free()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = runner.step(ddb, kp, rp, ep) if runner else ctx.assemble_iteration(ddb, kp, rp, ep)[0]
        tim.append(ctx.timings() if runner is None else runner.timings())
        n_out = out.n
        out.free()
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    ms_per_step = dt * 1000.0 / args.steps
    value = args.reads / (dt / args.steps)

    # ---- N > 1: the sharded result must equal the single-GPU result (checksums over hits, alignments, new DB) -------
    shard_check = None
    if runner is not None:
        shard_check = runner.verify_against_single_gpu(ddb, kp, rp, ep)
        if rank == 0:
            log("[bench/shard-check] %s" % json.dumps(shard_check))
            if not shard_check["equal"] and not os.environ.get("PLASS_B200_SHARD_CHECK_SOFT"):
                raise SystemExit("SHARDED RESULT DIFFERS from the single-GPU result: %s" % json.dumps(shard_check))

    if os.environ.get("PLASS_B200_BENCH_STOP_AFTER_CHECK"):
        if rank == 0:
            emit_json({"value": value, "ms_per_step": ms_per_step, "n_gpus": world, "stage_ms": {k: float(np.mean([t[k] for t in tim])) for k in tim[0] if k.endswith("_ms")},
                       "shard_check": shard_check})
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------------
    if runner is None:
        host = ddb.download()
        pinned = pin(host)
        del host
    else:
        pinned = runner.pinned_slice(ddb)                    # this rank's slice of the sequences, in pinned host memory
    h2d = int(pinned.data.nbytes + pinned.keys.nbytes + pinned.offsets.nbytes + pinned.lens.nbytes)
    e2e_times, d2h, e2e_phases = [], 0, []
    for i in range(3):
        barrier()
        t0 = time.perf_counter()
        if runner:
            d_in = runner.upload_sliced(pinned)              # PCIe: this rank's slice; NVLink: the rest
            out = runner.step(d_in, kp, rp, ep, download=True)
            d2h = runner.last_d2h_bytes
        else:
            d_in = ctx.upload(pinned)
            t1 = time.perf_counter()
            out, hits, alns = ctx.assemble_iteration(d_in, kp, rp, ep, want_intermediates=True)
            t2 = time.perf_counter()
            host_out = out.download()
            if i > 0:
                e2e_phases.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3))
            d2h = int(hits.nbytes + alns.nbytes + host_out.data.nbytes + host_out.offsets.nbytes + host_out.lens.nbytes + host_out.keys.nbytes)
            # the step's results live in pinned blocks of the library's pool: drop them so that the next step reuses them
            del hits, alns, host_out
        out.free(); d_in.free()
        barrier()
        if i > 0:
            e2e_times.append(time.perf_counter() - t0)
    e2e_dt = float(np.mean(e2e_times))
    e2e_serial_ms = e2e_dt * 1000.0
    e2e_mode = "one step at a time: upload, iteration, download of hits / alignments / new DB, all awaited before the next step"
    if runner is None:
        # Pipelined steps, the way the assembly workflow runs: the results of step i (prefilter hits, alignments, new DB)
        # travel to the host while step i+1 is uploaded and computed.  Every step still copies its input from pinned host
        # memory and every result is awaited inside the timed region (the last one after the loop).
        ddb.free(); ddb = None                                # the resident copy is not needed any more: room for two steps in flight
        ctx.set_async_results(True)
        try:
            trace = os.environ.get("PLASS_B200_E2E_TRACE")

            def run_pipelined(k):
                pending = None
                nbytes = 0
                nxt = ctx.upload_async(pinned)
                for i in range(k):
                    ta = time.perf_counter()
                    d_in = nxt
                    nxt = ctx.upload_async(pinned) if i + 1 < k else None
                    tb = time.perf_counter()
                    out, hits, alns = ctx.assemble_iteration(d_in, kp, rp, ep, want_intermediates=True)
                    tc = time.perf_counter()
                    host_out = out.download()
                    ticket = ctx.results_ticket()
                    out.free(); d_in.free()
                    td = time.perf_counter()
                    if pending is not None:
                        ctx.results_wait(pending[0])
                        nbytes = sum(int(a.nbytes) for a in pending[1:])
                    pending = (ticket, hits, alns, host_out.data, host_out.offsets, host_out.lens, host_out.keys)
                    if trace:
                        log("[bench/e2e] step %d: enqueue upload %.2f ms, iteration %.2f ms, enqueue download %.2f ms, wait previous results %.2f ms"
                            % (i, (tb - ta) * 1e3, (tc - tb) * 1e3, (td - tc) * 1e3, (time.perf_counter() - td) * 1e3))
                ctx.results_wait(pending[0])
                nbytes = sum(int(a.nbytes) for a in pending[1:])
                return nbytes
            run_pipelined(2)                                   # warm-up: two sets of pinned result blocks
            k = args.e2e_steps or max(args.steps, 8)           # long enough that the first upload and the final drain are amortised
            barrier()
            t0 = time.perf_counter()
            d2h = run_pipelined(k)
            barrier()
            e2e_dt = (time.perf_counter() - t0) / k
            e2e_mode = ("pipelined over %d steps: every step's input is copied from pinned host memory (upload stream) while the previous step computes, "
                        "the device-to-host copies of step i run under the kernels of step i+1; first upload and final drain are inside the timed region" % k)
        finally:
            ctx.set_async_results(False)
    else:
        # the same pipeline on every rank: this rank's slice of the next input travels over PCIe while the current step
        # computes; the slices are all-gathered over NVLink; this rank's share of the results goes back under the next step
        ctx.set_async_results(True)
        try:
            def run_pipelined_sharded(k):
                pending, nbytes = None, 0
                nxt = ctx.upload_async(pinned)
                for i in range(k):
                    sl_in = nxt
                    nxt = ctx.upload_async(pinned) if i + 1 < k else None
                    d_in = ctx.shard_allgather_db(sl_in)
                    sl_in.free()
                    out, _, hits, alns = ctx.shard_iteration(d_in, kp, rp, ep, want_intermediates=True)
                    host_out = out.download()
                    ticket = ctx.results_ticket()
                    out.free(); d_in.free()
                    if pending is not None:
                        ctx.results_wait(pending[0])
                        nbytes = sum(int(a.nbytes) for a in pending[1:])
                    pending = (ticket, hits, alns, host_out.data, host_out.offsets, host_out.lens, host_out.keys)
                ctx.results_wait(pending[0])
                return sum(int(a.nbytes) for a in pending[1:])
            run_pipelined_sharded(2)
            k = args.e2e_steps or max(args.steps, 8)
            barrier()
            t0 = time.perf_counter()
            d2h = run_pipelined_sharded(k)
            barrier()
            e2e_dt = (time.perf_counter() - t0) / k
            e2e_mode = ("pipelined over %d steps on every rank: the rank's slice of the next input is copied from pinned host memory while the current step computes, "
                        "the slices are all-gathered over NVLink, the rank's share of hits / alignments / new DB returns under the next step" % k)
        finally:
            ctx.set_async_results(False)
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
        t = torch.tensor([d2h, h2d], dtype=torch.int64, device="cuda")      # bytes of the whole job: sum over the ranks
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        d2h, h2d = int(t[0].item()), int(t[1].item())
    e2e_value = args.reads / e2e_dt

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    scatter_ms = float(np.mean([t["sort1_scatter_ms"] for t in tim]))
    passes = int(tim[-1]["sort1_passes"])
    nrec = int(tim[-1]["n_kmer_records"])
    alg_bytes = nrec * 16 * 2                      # a sort = one read + one write of every record
    achieved = alg_bytes / 1e9 / (scatter_ms / 1e3) if scatter_ms > 0 else 0.0
    traffic, traffic_src = roofline_traffic(nrec)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "radix scatter (sort #1, %d passes/step)" % passes,
                "algorithmic_bytes_per_launch": alg_bytes / max(passes, 1), "launch_ms": scatter_ms / max(passes, 1), "peak_source": peak_src}
    stage_keys = ("extract_ms", "sort1_ms", "group_ms", "sort2_ms", "reduce_ms", "rescore_ms", "extend_ms", "exchange_ms", "total_ms")
    stage_ms = {k: float(np.mean([t[k] for t in tim])) for k in stage_keys}
    if runner is not None and "exchange1_ms" in tim[-1]:
        # rank 0's two fused partition + exchange kernels (stream-ordered barrier included) and the bytes it sent over NVLink
        for k in ("exchange1_ms", "exchange2_ms"):
            stage_ms[k] = float(np.mean([t[k] for t in tim]))
        stage_ms["exchange1_bytes_sent"] = int(tim[-1]["exchange1_bytes_sent"])
        stage_ms["exchange2_bytes_sent"] = int(tim[-1]["exchange2_bytes_sent"])
    # per-stage position against the HBM roofline: algorithmic bytes of DESIGN.md section 3 / stage time / peak
    npair, nhit, naln = int(tim[-1]["n_pair_records"]), int(tim[-1]["n_hits"]), int(tim[-1]["n_alns"])
    stage_frac = None
    if world == 1:
        L = 48.0
        alg = {"extract_ms": n_frag_total * (L + 2) + 16.0 * nrec, "sort1_ms": 32.0 * nrec, "group_ms": 16.0 * nrec + 16.0 * npair,
               "sort2_ms": 32.0 * npair, "reduce_ms": 16.0 * npair + 16.0 * nhit, "rescore_ms": nhit * (L + 26) + n_frag_total * (L + 2),
               "extend_ms": naln * 24.0 + nhit * (L + 2) + 2 * n_frag_total * (L + 2)}
        stage_frac = {k: (alg[k] / 1e9 / (stage_ms[k] / 1e3) / peak if stage_ms[k] > 0 else None) for k in alg}
        stage_frac["iteration"] = sum(alg.values()) / 1e9 / (stage_ms["total_ms"] / 1e3) / peak

    # ---- BASELINE.json configs[2] / configs[3]: --num-iterations 3, chained (INPUT=assembly_$STEP, data/assemble.sh:153), with the
    # per-iteration parameters of src/workflow/Assembler.cpp:99-110 (hash shift 67, 68, 68; --include-only-extendable from
    # iteration 1).  N > 1: every iteration ends with the all-gather of the ranks' slices.  Timed for the record.
    chained = None
    if not args.no_extras:
        try:
            if ddb is None:
                ddb = ctx.upload(pinned)
            cur, per_it, t_all = ddb, [], time.perf_counter()
            for it in range(3):
                kpi = api.default_km_params(False, hash_shift=67 + (it + 1) // 2, include_only_extendable=1 if it > 0 else 0)
                barrier()
                t0 = time.perf_counter()
                if runner:
                    sl = runner.step(cur, kpi, rp, ep)
                    nxt = ctx.shard_allgather_db(sl)
                    sl.free()
                else:
                    nxt = ctx.assemble_iteration(cur, kpi, rp, ep)[0]
                barrier()
                per_it.append((time.perf_counter() - t0) * 1e3)
                if cur is not ddb:
                    cur.free()
                cur = nxt
            n_final = cur.n
            if cur is not ddb:
                cur.free()
            if world > 1:
                t = torch.tensor(per_it, dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                per_it = [float(x) for x in t.tolist()]
            chained = {"iterations": 3, "ms_per_iteration": per_it, "total_ms": float(sum(per_it)), "reads_per_s": args.reads * 3 / (sum(per_it) / 1e3),
                       "sequences_after": int(n_final), "note": "wall clock incl. the all-gather of the new DB at N > 1; hash shift 67, 68, 68; include-only-extendable 0, 1, 1"}
        except Exception as e:  # noqa: BLE001
            chained = {"failed": str(e)}

    # ---- BASELINE.json configs[4] at the size that fits: penguin nuclassemble, sharded, chained, cyclecheck in the loop --------
    nucl_sharded = None
    if runner is not None and not args.no_extras:
        try:
            n_nt = min(args.reads, 20000000)
            nkp, nrp, nep = api.default_km_params(True), api.default_rs_params(True), api.default_ex_params(True)
            dn = runner.build_and_broadcast(lambda: ctx.upload(synth.nucleotide_db(synth.make_reads_fast(n_nt, seed=args.seed + 100))))
            per_it, cyc = [], []
            cur = dn
            for it in range(2):
                barrier()
                t0 = time.perf_counter()
                sl = runner.step(cur, nkp, nrp, nep)
                nxt = ctx.shard_allgather_db(sl)
                barrier()
                per_it.append((time.perf_counter() - t0) * 1e3)
                t0 = time.perf_counter()
                split = ctx.cyclecheck(sl, 200000)            # every rank checks the contigs it owns (data/nuclassemble.sh:19-60)
                barrier()
                cyc.append((time.perf_counter() - t0) * 1e3)
                sl.free(); cur.free()
                cur = nxt
            cur.free()
            t = torch.tensor(per_it + cyc, dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            v = [float(x) for x in t.tolist()]
            nucl_sharded = {"reads": n_nt, "ranks": world, "iterations": 2, "ms_per_iteration": v[:2], "cyclecheck_ms": v[2:],
                            "reads_per_s": n_nt * 2 / ((v[0] + v[1]) / 1e3), "note": "k = 22, --min-seq-id 0.99, chained through the all-gathered DB"}
        except Exception as e:  # noqa: BLE001
            nucl_sharded = {"failed": str(e)}

    # ---- neighbouring steps of the iteration, measured for the record; not part of the metric ----------
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            extras = {}
            small = min(args.reads, 5000000)
            reads = synth.make_reads_fast(small, seed=args.seed)
            dn = ctx.upload(synth.nucleotide_db(reads))
            del reads
            ms = 0.0
            for rep_i in range(2):
                ms = 0.0
                lo, _ = ctx.extractorfs(dn, api.orf_params_long(), translate=True, want_info=False); ms += ctx.timings()["total_ms"]
                st, _ = ctx.extractorfs(dn, api.orf_params_start(), translate=True, want_info=False); ms += ctx.timings()["total_ms"]
                cat = ctx.concat(lo, st); ms += ctx.timings()["total_ms"]
                lo.free(); st.free()
                if rep_i == 0:
                    cat.free()
            extras["six_frame_fragments"] = {"reads": small, "fragments": int(cat.n), "ms": ms, "reads_per_s": small / (ms / 1e3)}
            # configs[1]: the 5 M-read iteration, and plass STEP 0 fused (two kmermatcher + rescorediagonal passes around findassemblystart)
            for _ in range(3):
                o5 = ctx.assemble_iteration(cat, kp, rp, ep)[0]
                t5 = ctx.timings()
                o5.free()
            extras["iteration_5M_reads"] = {"reads": small, "ms": t5["total_ms"], "reads_per_s": small / (t5["total_ms"] / 1e3), "stage_ms": {k: t5[k] for k in stage_keys[:7]}}
            for _ in range(2):
                corr, out0, _, _ = ctx.assemble_step0(cat, kp, rp, ep)
                t0s = ctx.timings()
                corr.free(); out0.free()
            extras["step0_fused_ms"] = t0s["total_ms"]
            cat.free()
            # the nucleotide path (penguin nuclassemble, k = 22): one iteration + cyclecheck on the same reads
            nkp, nrp, nep = api.default_km_params(True), api.default_rs_params(True), api.default_ex_params(True)
            for _ in range(2):
                nout = ctx.assemble_iteration(dn, nkp, nrp, nep)[0]
                tn = ctx.timings()
                n_nt_out = nout.n
                for _c in range(2):
                    split = ctx.cyclecheck(nout, 200000)
                    tc = ctx.timings()
                nout.free()
            dn.free()
            extras["nucl_iteration"] = {"reads": small, "ms": tn["total_ms"], "reads_per_s": small / (tn["total_ms"] / 1e3),
                                        "stage_ms": {k: tn[k] for k in stage_keys[:7]},
                                        "kmer_records": int(tn["n_kmer_records"]), "hits": int(tn["n_hits"]), "alignments": int(tn["n_alns"]),
                                        "output_sequences": int(n_nt_out)}
            extras["cyclecheck"] = {"sequences": int(n_nt_out), "ms": tc["total_ms"], "reported": int((split > 0).sum())}
        except Exception as e:  # noqa: BLE001
            extras = dict(extras or {}, failed=str(e))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32 (+f64 E-values)", "data": "synthetic",
            "config": workload_config(args.reads),
            "workload_stats": {"fragments": int(n_frag_total), "kmer_records": nrec, "pair_records": npair, "hits": nhit, "alignments": naln,
                               "output_sequences": int(n_out), "kmer_record_bytes_per_step": nrec * 16,
                               "parallelism": (runner.describe() if runner else "single GPU"),
                               "counts_are": "rank 0's share" if world > 1 else "the whole job's"},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_dt * 1000.0,
                    "mode": e2e_mode, "serial_ms_per_step": e2e_serial_ms,
                    "phases_ms": ({"upload": float(np.mean([p[0] for p in e2e_phases])), "iteration_with_hits_alns_d2h": float(np.mean([p[1] for p in e2e_phases])),
                                   "download_new_db": float(np.mean([p[2] for p in e2e_phases]))} if e2e_phases else (runner.e2e_phases() if runner else None))},
            "gpu_launches": int(sum(t["kernel_launches"] for t in tim)),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "dropin": dropin, "shard_check": shard_check,
            "clocks": clocks, "stage_ms": stage_ms, "stage_roofline_frac": stage_frac, "three_iterations": chained, "nucl_sharded": nucl_sharded, "extras": extras,
        }
        emit_json(line)
    if ddb is not None:
        ddb.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
