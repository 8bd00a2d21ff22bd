#!/usr/bin/env python3
"""bench.py -- reads/s through one full assemble iteration (kmermatcher + rescorediagonal + assembleresults).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--reads R]

Workload (BASELINE.json configs[1]): R = 5 M synthetic 150 bp protein-coding reads (seed 1, 20x coverage,
0.5 % substitutions, 50 % reverse-complemented) -> six-frame amino-acid fragments; one `plass assemble`
iteration with the workflow defaults (k = 14, 13-letter alphabet, --min-seq-id 0.9, -e 1e-5).
A "step" is one pass of the hot path over that fragment DB.

  value      whole-job reads/s with the fragment DB already resident in HBM (device-resident fused iteration)
  e2e        the same iteration through the C ABI with HOST buffers: pinned H2D of the DB, D2H of the
             prefilter hits, the alignments and the new sequence DB inside the timed region
  roofline   dominant kernel = the radix scatter passes of sort #1; achieved = algorithmic bytes of the sort
             (one read + one write of every 16-byte record) / summed scatter time
  cpu_baseline  the UNMODIFIED reference binary (oracle/_ref/bin/plass: kmermatcher, rescorediagonal,
             assembleresults sub-commands, all host threads) on a bounded sample of the same fragments

--impl reference times only that reference arm (CPU) and prints it as the line's value.
N > 1 (torchrun): weak scaling -- R reads per GPU; extraction is sliced by sequence, the k-mer records go to the
rank owning the k-mer and the candidate pairs to the rank owning the representative (two NCCL all-to-alls through
torch.distributed), see DESIGN.md §5.
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

from plass_b200 import api, mmseqsdb, synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "plass")
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


def workload_name(reads_per_gpu, per_gpu):
    cfg = {5000000: " (BASELINE configs[1])", 50000000: " (BASELINE configs[2], one iteration)"}.get(reads_per_gpu, "")
    return "%gM synthetic 150bp coding reads%s -> aa fragments, k=14, alph 13, --min-seq-id 0.9, 1 iteration%s" % (
        reads_per_gpu / 1e6, " per GPU" if per_gpu else "", cfg)


def make_fragments(n_reads, seed):
    cache = os.path.join(tempfile.gettempdir(), "plass_b200_frag_%d_%d.npz" % (n_reads, seed))
    if os.path.exists(cache):
        z = np.load(cache)
        return mmseqsdb.DB(z["data"], z["keys"], z["offsets"], z["lens"], 0)
    t = time.time()
    reads = synth.make_reads_fast(n_reads, seed=seed)
    db = synth.protein_fragments(reads, workers=min(16, os.cpu_count() or 1))
    del reads
    log("[bench] generated %d reads -> %d aa fragments (mean %.1f aa) in %.1f s" % (n_reads, db.n, float(db.lens.mean()) - 2, time.time() - t))
    if n_reads <= 45000000:      # the cache only serves the other ranks of a multi-GPU run (up to 8 x 5 M reads) and repeated small samples
        try:
            np.savez(cache, data=db.data, keys=db.keys, offsets=db.offsets, lens=db.lens)
        except OSError:
            pass
    return db


def subsample(db, n_frag):
    """First n_frag fragments as their own DB (keys renumbered 0..n-1)."""
    n = min(n_frag, db.n)
    end = int(db.offsets[n - 1]) + int(db.lens[n - 1])
    return mmseqsdb.DB(db.data[:end], np.arange(n, dtype=np.uint32), db.offsets[:n].copy(), db.lens[:n].copy(), db.dbtype)


def write_db_fast(path, db):
    db.data.tofile(path)
    with open(path + ".index", "w") as f:
        f.write("".join("%d\t%d\t%d\n" % (int(k), int(o), int(l)) for k, o, l in zip(db.keys, db.offsets, db.lens)))
    np.array([db.dbtype], dtype="<i4").tofile(path + ".dbtype")


def run_reference_iteration(db, threads, workdir):
    """kmermatcher + rescorediagonal + assembleresults of the unmodified reference on `db`; returns seconds."""
    if os.path.exists(workdir):
        shutil.rmtree(workdir)
    os.makedirs(workdir)
    seq = os.path.join(workdir, "seq")
    write_db_fast(seq, db)
    env = dict(os.environ, MMSEQS_NUM_THREADS=str(threads))
    cmds = [
        [REF_BIN, "kmermatcher", seq, os.path.join(workdir, "pref")] + KM_FLAGS.split() + ["--threads", str(threads)],
        [REF_BIN, "rescorediagonal", seq, seq, os.path.join(workdir, "pref"), os.path.join(workdir, "aln")] + RS_FLAGS.split() + ["--threads", str(threads)],
        [REF_BIN, "assembleresults", seq, os.path.join(workdir, "aln"), os.path.join(workdir, "asm")] + EX_FLAGS.split() + ["--threads", str(threads)],
    ]
    total = 0.0
    per = []
    for c in cmds:
        t = time.time()
        r = subprocess.run(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
        dt = time.time() - t
        if r.returncode != 0:
            raise RuntimeError("reference step failed: %s\n%s" % (" ".join(c[:2]), r.stdout.decode()[-2000:]))
        per.append(dt)
        total += dt
    return total, per


def sized_cpu_sample(args, n_reads_total, threads, work):
    """Bounded sample for the CPU arm: the same generator at the same 20x coverage (a prefix of the big DB would have
    lower coverage, fewer overlaps per read, and flatter the CPU).  A first run on --cpu-sample-reads reads calibrates;
    if it took less than ~10 s the sample is enlarged towards ~15 s of CPU work, up to the whole workload."""
    sample_reads = min(args.cpu_sample_reads, n_reads_total)
    sample = make_fragments(sample_reads, args.seed)
    t, per = run_reference_iteration(sample, threads, work)
    log("[bench/reference] calibration: %d reads in %.2f s" % (sample_reads, t))
    if t < 10.0 and sample_reads < n_reads_total:
        want = int(sample_reads * 15.0 / max(t, 0.05))
        want = min(n_reads_total, max(sample_reads, (want // 100000) * 100000))
        if want > sample_reads:
            sample_reads = want
            sample = make_fragments(sample_reads, args.seed)
    return sample_reads, sample


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


def reference_arm(args, db_full, n_reads_total):
    """--impl reference: the reference's own CPU path, bounded sample, all host threads."""
    threads = os.cpu_count() or 1
    # same generator, same 20x coverage, fewer reads (a prefix of the big DB would have lower coverage and
    # therefore fewer overlaps per read, which would flatter the CPU)
    work = os.path.join(tempfile.gettempdir(), "plass_b200_ref_%d" % os.getpid())
    sample_reads, sample = sized_cpu_sample(args, n_reads_total, threads, work)
    times = []
    try:
        for i in range(args.warmup + args.steps):
            t, per = run_reference_iteration(sample, threads, work)
            log("[bench/reference] step %d: %.2f s (kmermatcher %.2f, rescorediagonal %.2f, assembleresults %.2f)" % (i, t, per[0], per[1], per[2]))
            if i >= args.warmup:
                times.append(t)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    sec = float(np.mean(times))
    val = sample_reads / sec
    sample_desc = "%d reads (%d fragments) from the same generator at the same 20x coverage instead of %d reads; %d threads; tmp dir %s" % (
        sample_reads, sample.n, n_reads_total, threads, tempfile.gettempdir())
    return {
        "metric": "reads/sec per assemble iteration (kmermatcher+rescorediagonal+assembleresults)", "value": val, "unit": "reads/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1000.0,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (+f64 E-values)", "data": "synthetic",
        "config": {"workload": workload_name(args.reads, False) + "; bounded sample",
                   "reads": n_reads_total, "fragments": int(db_full.n)},
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
    ap.add_argument("--reads", type=int, default=5000000, help="reads per GPU")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample-reads", type=int, default=400000, help="reads in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the (untimed-for-the-metric) measurements of the neighbouring steps")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        db = make_fragments(args.reads, args.seed)
        emit_json(reference_arm(args, db, args.reads))
        return 0

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)

    # weak scaling: every rank contributes `reads` reads; the job's DB is the union (replicated in each HBM)
    n_reads_total = args.reads * world
    if world > 1:
        # rank 0 generates (and caches) the job's input, the other ranks load the cached copy
        db = make_fragments(n_reads_total, args.seed) if rank == 0 else None
        dist.barrier()
        if db is None:
            db = make_fragments(n_reads_total, args.seed)
    else:
        db = make_fragments(n_reads_total, args.seed)
    ctx = api.Context(local_rank)
    kp, rp, ep = api.default_km_params(False), api.default_rs_params(False), api.default_ex_params(False)

    if world > 1:
        from plass_b200 import sharded
        runner = sharded.ShardedIteration(ctx, dist, rank, world)
    else:
        runner = None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident arm (value) -------------------------------------------------------------
    ddb = ctx.upload(db)
    peak, peak_src = load_peaks()
    tim = []
    # the clock sampler (nvidia-smi, one line every 50 ms) is started before the warm-up steps: the tool needs ~0.1 s
    # before its first line, and the timed region of a few 50 ms steps would otherwise see a single sample.  Warm-up and
    # timed steps run the same kernels, so every sample is taken under the load that is being measured.
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
    value = n_reads_total / (dt / args.steps)

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------------
    pinned = mmseqsdb.DB(torch.from_numpy(db.data).pin_memory().numpy(), torch.from_numpy(db.keys).pin_memory().numpy(),
                         torch.from_numpy(db.offsets.view(np.int64)).pin_memory().numpy().view(np.uint64),
                         torch.from_numpy(db.lens.view(np.int32)).pin_memory().numpy().view(np.uint32), db.dbtype)
    h2d = int(pinned.data.nbytes + pinned.keys.nbytes + pinned.offsets.nbytes + pinned.lens.nbytes)
    e2e_times, d2h, e2e_phases = [], 0, []
    e2e_steps = max(1, min(args.steps, 2))
    for i in range(1 + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        if runner:
            d_in, h2d_rank = sharded.upload_sliced(ctx, dist, pinned, rank, world)   # PCIe: this rank's slice; NVLink: the rest
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
            # the step's results live in pinned blocks of the library's pool: drop them so that the next step reuses
            # the blocks instead of pinning 2 GB of fresh host memory
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
        ctx.set_async_results(True)
        try:
            trace = os.environ.get("PLASS_B200_E2E_TRACE")

            def run_pipelined(k):
                pending = None
                nbytes = 0
                # every step copies its own input from pinned host memory; the copy of step i+1's input is enqueued on
                # the upload stream before step i's kernels are launched, so it travels underneath them
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
            k = max(args.steps, 8)                             # long enough that the first upload and the final drain are amortised
            barrier()
            t0 = time.perf_counter()
            d2h = run_pipelined(k)
            barrier()
            e2e_dt = (time.perf_counter() - t0) / k
            e2e_mode = ("pipelined over %d steps: every step's input is copied from pinned host memory (upload stream) while the previous step computes, "
                        "the device-to-host copies of step i run under the kernels of step i+1; first upload and final drain are inside the timed region" % k)
        finally:
            ctx.set_async_results(False)
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
        t = torch.tensor([d2h], dtype=torch.int64, device="cuda")      # bytes of the whole job: sum over the ranks
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        d2h = int(t.item())
    e2e_value = n_reads_total / e2e_dt

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    scatter_ms = float(np.mean([t["sort1_scatter_ms"] for t in tim]))
    passes = int(tim[-1]["sort1_passes"])
    nrec = int(tim[-1]["n_kmer_records"])
    alg_bytes = nrec * 16 * 2                      # a sort = one read + one write of every record
    achieved = alg_bytes / 1e9 / (scatter_ms / 1e3) if scatter_ms > 0 else 0.0
    # DRAM traffic of one scatter launch from the ncu --set full capture of this kernel (profiles/r1_summary_c.md):
    # 8.14 GB for 250 M records = 32.56 B per record and pass, i.e. the algorithmic 2 x 16 B plus 1.7 % (status words, tails)
    traffic = nrec * 32.56
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "radix_scatter_kernel (sort #1, %d launches/step)" % passes,
                "algorithmic_bytes_per_launch": alg_bytes / max(passes, 1), "launch_ms": scatter_ms / max(passes, 1), "peak_source": peak_src}
    stage_ms = {k: float(np.mean([t[k] for t in tim])) for k in ("extract_ms", "sort1_ms", "group_ms", "sort2_ms", "reduce_ms", "rescore_ms", "extend_ms", "exchange_ms", "total_ms")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        try:
            threads = os.cpu_count() or 1
            work = os.path.join(tempfile.gettempdir(), "plass_b200_cpu_%d" % os.getpid())
            sample_reads, sample = sized_cpu_sample(args, n_reads_total, threads, work)
            t, per = run_reference_iteration(sample, threads, work)
            shutil.rmtree(work, ignore_errors=True)
            cpu = {"value": sample_reads / t, "unit": "reads/s", "cores": threads, "kind": "reference",
                   "sample": "%d reads (%d fragments), same generator and 20x coverage, one iteration, %.1f s (kmermatcher %.1f, rescorediagonal %.1f, assembleresults %.1f)"
                             % (sample_reads, sample.n, t, per[0], per[1], per[2])}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}

    # ---- neighbouring steps of the iteration, measured for the record; not part of the metric ----------
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            extras = {}
            # plass STEP 0 fused: two kmermatcher + rescorediagonal passes around findassemblystart, then assembleresults
            for _ in range(2):
                corr, out0, _, _ = ctx.assemble_step0(ddb, kp, rp, ep)
                t0 = ctx.timings()
                corr.free(); out0.free()
            extras["step0_fused_ms"] = t0["total_ms"]
            # reads -> aa_6f_start_long on the GPU (extractorfs x 2 + translatenucs x 2 + concatdbs, data/assemble.sh:41-77)
            reads = synth.make_reads_fast(min(args.reads, 5000000), seed=args.seed)
            dn = ctx.upload(synth.nucleotide_db(reads))
            n_reads_orf = int(reads.shape[0])
            del reads
            ms = 0.0
            for rep_i in range(2):
                ms = 0.0
                lo, _ = ctx.extractorfs(dn, api.orf_params_long(), translate=True, want_info=False); ms += ctx.timings()["total_ms"]
                st, _ = ctx.extractorfs(dn, api.orf_params_start(), translate=True, want_info=False); ms += ctx.timings()["total_ms"]
                cat = ctx.concat(lo, st); ms += ctx.timings()["total_ms"]
                n_frag = cat.n
                lo.free(); st.free(); cat.free()
            dn.free()
            extras["six_frame_fragments_ms"] = ms
            extras["six_frame_fragments"] = {"reads": n_reads_orf, "fragments": int(n_frag), "reads_per_s": n_reads_orf / (ms / 1e3)}
            # the nucleotide path (penguin nuclassemble, k = 22): one iteration + cyclecheck on the same reads
            nreads = synth.make_reads_fast(min(args.reads, 5000000), seed=args.seed + 100)
            dn = ctx.upload(synth.nucleotide_db(nreads))
            del nreads
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
            extras["nucl_iteration"] = {"reads": int(min(args.reads, 5000000)), "ms": tn["total_ms"],
                                        "stage_ms": {k: tn[k] for k in ("extract_ms", "sort1_ms", "group_ms", "sort2_ms", "reduce_ms", "rescore_ms", "extend_ms")},
                                        "kmer_records": int(tn["n_kmer_records"]), "hits": int(tn["n_hits"]), "alignments": int(tn["n_alns"]),
                                        "output_sequences": int(n_nt_out)}
            extras["cyclecheck"] = {"sequences": int(n_nt_out), "ms": tc["total_ms"], "reported": int((split > 0).sum())}
        except Exception as e:  # noqa: BLE001
            extras = dict(extras or {}, failed=str(e))

    if rank == 0:
        line = {
            "metric": "reads/sec per assemble iteration (kmermatcher+rescorediagonal+assembleresults)", "value": value, "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (+f64 E-values)", "data": "synthetic",
            "config": {"workload": workload_name(args.reads, True),
                       "reads": n_reads_total, "fragments": int(db.n), "kmer_records": nrec, "pair_records": int(tim[-1]["n_pair_records"]),
                       "hits": int(tim[-1]["n_hits"]), "alignments": int(tim[-1]["n_alns"]), "output_sequences": int(n_out),
                       "l2": "inputs_larger_than_l2 (%.1f GB of k-mer records per step)" % (nrec * 16 / 1e9),
                       "parallelism": ("%d ranks: sequence-sliced extraction, all-to-all of k-mer records by k-mer owner, all-to-all of pair records by "
                                       "representative owner; record counts above are rank 0's share" % world) if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_dt * 1000.0,
                    "mode": e2e_mode, "serial_ms_per_step": e2e_serial_ms,
                    "phases_ms": ({"upload": float(np.mean([p[0] for p in e2e_phases])), "iteration_with_hits_alns_d2h": float(np.mean([p[1] for p in e2e_phases])),
                                   "download_new_db": float(np.mean([p[2] for p in e2e_phases]))} if e2e_phases else None)},
            "gpu_launches": int(sum(t["kernel_launches"] for t in tim)),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "stage_ms": stage_ms, "extras": extras,
        }
        emit_json(line)
    ddb.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
