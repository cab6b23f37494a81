#!/usr/bin/env python
"""bench.py -- throughput of the `abismal map` hot path on B200.

Workloads (--mode), all on a synthetic i.i.d. 3.1 Gbp genome with reads from the reference's own `sim`:
  pbat   150 bp paired-end PBAT reads (`sim -a`), mapped with -P      BASELINE configs[3]  (default, the metric)
  rpbat  150 bp paired-end random-PBAT reads (`sim -R`), mapped -R    BASELINE configs[4]
  se     150 bp single-end reads (`sim -single`), default mode        BASELINE configs[2]
A "step" maps one batch per GPU (--pairs read pairs, or 2 x --pairs single-end reads); reads are sharded over
GPUs with the index replicated (weak scaling, no data-path collective; NCCL only sums the mapping statistics).

  value  reads/s with the batch already resident in HBM (CUDA-event time of the K launches, max over ranks)
  e2e    reads/s through the public call (abg_map_batch: host buffers in, host buffers out, H2D/D2H inside
         the timed region)
  fastq_to_sam   (N = 1) reads/s of the whole front end, `abismal-b200 map` FASTQ -> SAM, i.e. what the
         reference arm times
  parity every rank checks the records of the first pairs of its timed batch against the CPU oracle; rank 0
         (N = 1) also compares the SAM of `abismal-b200 map` with the SAM of the unmodified reference binary
         on the cpu_baseline sample
  --impl reference   the unmodified reference binary (oracle/_ref/abismal map -t <all cores>) on bounded
         samples of the same reads and index.

A rank that fails writes its traceback to stderr and to gpurun_out/bench_logs/rank<r>.log, and rank 0 prints
a JSON line carrying every rank's error before the process group exits non-zero.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "abismal")
CLI = os.path.join(ROOT, "abismal_b200", "bin", "abismal-b200")

MODES = {
    # sim flag, map flags, paired, metric, workload text
    "pbat": ("-a", ["-P"], True, "mapped reads/sec (150bp PE bisulfite, 3.1 Gbp genome)",
             "150bp PE PBAT reads (sim -a -m 0.01 -b 0.98, fragments 150-400), abismal map -P", "configs[3]"),
    "rpbat": ("-R", ["-R"], True, "mapped reads/sec (150bp PE random-PBAT bisulfite, 3.1 Gbp genome)",
              "150bp PE random-PBAT reads (sim -R -m 0.01 -b 0.98, fragments 150-400), abismal map -R", "configs[4]"),
    "se": (None, [], False, "mapped reads/sec (150bp SE bisulfite, 3.1 Gbp genome)",
           "150bp SE reads (sim -single -m 0.01 -b 0.98), abismal map", "configs[2]"),
}

RANK = int(os.environ.get("RANK", "0"))


def log(*a):
    if RANK == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


def rank_log(text):
    """Every rank: stderr and a file that survives torchrun's summary."""
    sys.stderr.write("[bench rank %d] %s\n" % (RANK, text))
    sys.stderr.flush()
    try:
        d = os.path.join(ROOT, "gpurun_out", "bench_logs")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "rank%d.log" % RANK), "a") as f:
            f.write(text + "\n")
    except OSError:
        pass


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for ln in self.proc.stdout:
                self.lines.append(ln.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sample_fastq(src, dst, first, count):
    """Records [first, first + count) of a 4-line-record FASTQ."""
    with open(src, "rb") as fi, open(dst, "wb") as fo:
        for k, ln in enumerate(fi):
            if k < 4 * first:
                continue
            if k >= 4 * (first + count):
                break
            fo.write(ln)


def _map_seconds(cmd):
    """Run a `map -v` command -> (mapping seconds = wall - index loading, loading seconds, stderr)."""
    t = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t
    if p.returncode != 0:
        raise RuntimeError("%s failed: %s" % (" ".join(cmd[:3]), p.stderr[-800:]))
    load = 0.0
    for ln in p.stderr.splitlines():
        if "loading time:" in ln:
            load = float(ln.split("loading time:")[1].strip().rstrip("s"))
    return wall - load, load, p.stderr


def run_reference_map(index_path, fqs, n_threads, flags, out_sam):
    return _map_seconds([REF_BIN, "map", "-v", "-t", str(n_threads)] + list(flags) +
                        ["-i", index_path, "-o", out_sam] + list(fqs))[:2]


def run_cli_map(index_path, fqs, flags, out_sam, extra=()):
    """`abismal-b200 map` FASTQ -> SAM.  -> (mapping seconds incl. the index upload to HBM, stage busy times text,
    seconds of that upload)"""
    secs, load, err = _map_seconds([CLI, "map", "-v"] + list(extra) + list(flags) + ["-i", index_path, "-o", out_sam] + list(fqs))
    stages, upload = "", 0.0
    for ln in err.splitlines():
        if "total mapping time:" in ln:
            secs = float(ln.split("total mapping time:")[1].strip().rstrip("s"))
        if "index upload to HBM:" in ln:
            upload = float(ln.split("index upload to HBM:")[1].split("s")[0])
        if "stage busy time:" in ln:
            stages = ln.split("stage busy time:")[1].strip()
    return secs, stages, upload


def sam_records_differing(a_path, b_path):
    """SAM bodies (no @PG) compared as sorted multisets: the reference with -t > 1 writes its 1000-read batches in
    the order its threads finish (Threading).  -> (records in a, records that differ)"""
    def body(p):
        with open(p, "rb") as f:
            return sorted(ln for ln in f if not ln.startswith(b"@PG"))
    a, b = body(a_path), body(b_path)
    if a == b:
        return len(a), 0
    from collections import Counter
    ca, cb = Counter(a), Counter(b)
    return len(a), sum(((ca - cb) + (cb - ca)).values())


def algorithmic_bytes(counters, units):
    """bytes = 8 N_lookup + 4 N_entry + 8 (N_word + N_cmp) + 0.5 N_dpref, per unit.
    -> (seeding part: counter probes, index entries, packed compares; alignment part: DP reference bases)"""
    c = counters
    return ((8 * c["n_lookup"] + 4 * c["n_entry"] + 8 * (c["n_word"] + c["n_cmp"])) / units,
            0.5 * c["n_dpref"] / units)


class RankSync:
    """Keeps the ranks in step through failures: after every phase all ranks learn whether any of them failed,
    so nobody is left waiting in a barrier for a rank that raised."""

    def __init__(self, dist, use_dist, world):
        self.dist, self.use_dist, self.world = dist, use_dist, world
        self.error = None

    def run(self, what, fn):
        """fn() on this rank unless an earlier phase failed somewhere; -> fn's value or None."""
        val = None
        if self.error is None:
            try:
                val = fn()
            except BaseException:  # incl. SystemExit from helpers
                self.error = "%s: %s" % (what, traceback.format_exc())
                rank_log("FAILED in " + self.error)
        self.check()
        return val

    def check(self):
        if not self.use_dist:
            if self.error is not None:
                self.finish()
            return
        errs = [None] * self.world
        self.dist.all_gather_object(errs, self.error)
        if any(e is not None for e in errs):
            self.finish(errs)

    def finish(self, errs=None):
        errs = errs if errs is not None else [self.error]
        if RANK == 0:
            print(json.dumps({"error": "bench failed", "ranks": {str(r): e for r, e in enumerate(errs) if e is not None}}))
            sys.stdout.flush()
        if self.use_dist:
            try:
                self.dist.destroy_process_group()
            except Exception:
                pass
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="pbat", choices=sorted(MODES))
    ap.add_argument("--genome-bases", type=float, default=3.1e9)
    ap.add_argument("--pairs", type=int, default=1 << 20, help="read pairs (or pairs of single-end reads) per GPU per step")
    ap.add_argument("--cpu-sample-pairs", type=int, default=200000,
                    help="pairs per run of the reference binary (cpu_baseline, and every step of --impl reference)")
    ap.add_argument("--parity-pairs", type=int, default=4000, help="pairs per rank checked against the CPU oracle")
    ap.add_argument("--seed", type=int, default=20251017)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--as-rank", type=int, default=None,
                    help="single process: time the reads rank R of a multi-GPU run would get (reproduces one rank of an N-GPU run)")
    ap.add_argument("--no-cli", action="store_true", help="skip the FASTQ->SAM leg and the SAM comparison with the reference")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and RANK != 0:
        return 0

    import torch  # plumbing: device selection, NCCL, barriers
    import torch.distributed as dist

    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device to prepare the synthetic index"}))
            return 0
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sync = RankSync(dist, use_dist, world)
    try:
        return run(args, sync, torch, dist, local_rank, world, use_dist)
    except SystemExit:
        raise
    except BaseException:
        rank_log("FAILED outside a phase: " + traceback.format_exc())
        raise


def run(args, sync, torch, dist, local_rank, world, use_dist):
    from abismal_b200 import workload
    rank = RANK
    sim_flag, map_flags, paired, metric, wl_text, cfg_name = MODES[args.mode]
    genome_bases = int(args.genome_bases)
    n_cpu = os.cpu_count() or 1
    n_units = args.pairs if paired else 2 * args.pairs   # work items (pairs / single-end reads) per GPU per step
    reads_per_unit = 2 if paired else 1
    unit = "pairs" if paired else "reads"

    # ---- workload (untimed): genome, index, reads --------------------------------
    state = {}

    def prep_index():
        if rank == 0:
            state["ixf"], state["paths"] = workload.get_index(genome_bases, args.seed, device=local_rank, need_files=True, log=log)
    sync.run("index preparation (rank 0)", prep_index)

    def prep_reads():
        if "ixf" not in state:
            state["ixf"], state["paths"] = workload.get_index(genome_bases, args.seed, device=local_rank, need_files=True, log=log)
        paths = state["paths"]
        sim_procs = int(os.environ.get("ABISMAL_B200_SIM_PROCS", "0")) or max(1, min(16, n_cpu // max(world, 1)))
        data_rank = rank if args.as_rank is None else args.as_rank
        prefix = os.path.join(paths["dir"], "%s_n%d_r%d" % (args.mode, n_units, data_rank))
        state["prefix"] = prefix
        fq1, fq2 = workload.simulate_reads(REF_BIN, paths["fasta"], prefix, n_units, seed=args.seed % 1000 + data_rank,
                                           paired=paired, mode_flag=sim_flag, n_procs=sim_procs, log=log)
        state["fqs"] = [fq1, fq2] if paired else [fq1]
    sync.run("read simulation", prep_reads)
    paths, prefix, fqs, ixf = state["paths"], state["prefix"], state["fqs"], state["ixf"]

    config = {
        "workload": "synthetic i.i.d. %.2f Gbp genome (24 chroms), %s; one batch of %d %s per GPU per step "
                    "(BASELINE %s shape)" % (genome_bases / 1e9, wl_text, n_units, unit, cfg_name),
        "%s_per_gpu_per_step" % unit: n_units,
        "index": "replicated per GPU, built on GPU (byte-identical to `abismal idx`)",
        "parallelism": "reads sharded x%d, no data-path collective" % world,
        "l2": "inputs larger than L2: %.1f GB index gathered at random + %.0f MB of reads per step"
              % (2.7 * genome_bases / 3.1e9, n_units * reads_per_unit * 150 / 1e6),
    }

    if args.impl == "reference":
        # ---- reference arm: unmodified reference binary, all host cores, bounded samples -------------------
        # every step maps the NEXT --cpu-sample-pairs units of the step batch (wrapping around), so K steps cover
        # K x sample distinct reads of the same workload
        n_s = min(args.cpu_sample_pairs, n_units)
        out_sam = prefix + "_ref_step.sam"
        secs = []
        for k in range(args.warmup + args.steps):
            first = (k * n_s) % max(1, n_units - n_s + 1)
            s = [prefix + "_refstep_%d.fq" % (e + 1) for e in range(len(fqs))]
            for src, dst in zip(fqs, s):
                sample_fastq(src, dst, first, n_s)
            t, _ = run_reference_map(paths["index"], s, n_cpu, map_flags, out_sam)
            if k >= args.warmup:
                secs.append(t)
        for f in [out_sam] + s:
            try:
                os.remove(f)
            except OSError:
                pass
        total = sum(secs)
        value = reads_per_unit * n_s * args.steps / total
        sample = ("%d %s of the step batch per step (a different slice every step, %d %s timed in all), abismal map -t %d %s "
                  "FASTQ -> SAM (mapping time = wall - index loading)" % (n_s, unit, n_s * args.steps, unit, n_cpu, " ".join(map_flags)))
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/u64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": n_cpu, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ---- our arm -----------------------------------------------------------------
    from abismal_b200 import Index, Mapper, MODE_A_RICH, MODE_PAIRED, MODE_RANDOM_PBAT
    from abismal_b200.capi import Results
    mode = {"pbat": MODE_PAIRED | MODE_A_RICH, "rpbat": MODE_PAIRED | MODE_RANDOM_PBAT, "se": 0}[args.mode]

    def setup():
        t = time.time()
        b = [workload.load_fastq_fast(f) for f in fqs]
        log("reads loaded in %.1fs" % (time.time() - t))
        t = time.time()
        ix = Index(ixf, local_rank)
        log("index resident in HBM: %.2f GB, uploaded in %.1fs, features %d" % (ix.device_bytes / 1e9, time.time() - t, ix.features))
        m = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max([x.max_len for x in b] + [64]))
        res = Results(b[0].n, paired, m.stride, pinned=True)
        b = [x.to_pinned() for x in b]  # the e2e leg copies each step's inputs from pinned host memory
        state.update(b=b, ix=ix, m=m, res=res)
    sync.run("setup (reads, index upload, mapper)", setup)
    b, ix, m, res = state["b"], state["ix"], state["m"], state["res"]

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # device-resident leg
    def warm():
        m.upload(*b)
        m.sync()
        for _ in range(args.warmup):
            m.run()
            m.sync()
    sync.run("upload + warm-up", warm)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()

    def timed_device():
        kernel_ms, phase_ms, seed_ms = [], [0.0] * 5, [0.0] * 4
        for _ in range(args.steps):
            m.run()
            m.sync()
            kernel_ms.append(m.last_kernel_ms)
            phase_ms = [x + y for x, y in zip(phase_ms, m.last_kernel_times)]
            seed_ms = [x + y for x, y in zip(seed_ms, m.last_seed_times)]
        state.update(dev_ms=sum(kernel_ms), phase_ms=phase_ms, seed_ms=seed_ms, bin_stats=m.bin_stats() if m.binned else None)
    sync.run("device-resident leg", timed_device)
    barrier()
    dev_ms, phase_ms = state["dev_ms"], state["phase_ms"]
    launches = args.steps * m.launches_per_run

    # end-to-end leg through the public call
    def warm_e2e():
        for _ in range(2):
            m.map_batch(*b, results=res)
    sync.run("e2e warm-up", warm_e2e)
    barrier()

    def timed_e2e():
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m.map_batch(*b, results=res)
        torch.cuda.synchronize()
        state["e2e_s"] = time.perf_counter() - t0
    sync.run("e2e leg", timed_e2e)
    barrier()
    e2e_s = state["e2e_s"]
    clocks = sampler.stop()

    # ---- parity of this rank's timed batch against the CPU oracle (checker only) -------------------------
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    n_o = min(args.parity_pairs, b[0].n)

    def parity():
        import helpers
        o = helpers.OracleMapper(ixf, mode=mode)
        t = time.perf_counter()
        want = o.map_batch(*[x.slice(0, n_o) for x in b])
        state["port_s"] = time.perf_counter() - t
        keys = ("se1", "n_cigar1") + (("pe_r1", "pe_r2", "se2", "n_cigar2") if paired else ())
        bad = 0
        for k in keys:
            bad += int((getattr(res, k)[:n_o] != getattr(want, k)[:n_o]).sum())
        # CIGAR operations of the checked records
        for e in (1, 2) if paired else (1,):
            cg, ng = (res.cigar1, res.n_cigar1) if e == 1 else (res.cigar2, res.n_cigar2)
            cw, nw = (want.cigar1, want.n_cigar1) if e == 1 else (want.cigar2, want.n_cigar2)
            for i in range(n_o):
                k = min(int(ng[i]), cg.shape[1])
                if int(ng[i]) == int(nw[i]) and not (cg[i, :k] == cw[i, :k]).all():
                    bad += 1
        state["bad"] = bad
        state["oracle_counters"] = o.counters.as_dict()
        o.close()
        if bad:
            raise RuntimeError("%d result records differ from the CPU oracle on the first %d %s of rank %d's batch"
                               % (bad, n_o, unit, rank))
    sync.run("parity against the CPU oracle", parity)

    mapped = int(((res.pe_r1 if paired else res.se1)["pos"] != 0).sum())
    stats = torch.tensor([dev_ms, e2e_s, float(b[0].n), float(mapped), float(state["bad"]), float(n_o)],
                         dtype=torch.float64, device="cuda")
    if use_dist:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)  # NCCL: the only collective (mapping statistics)
        dev_ms_max, e2e_max = float(mx[0]), float(mx[1])
        total_units, total_mapped, total_bad, total_checked = float(sm[2]), float(sm[3]), int(sm[4]), int(sm[5])
    else:
        dev_ms_max, e2e_max, total_units, total_mapped = dev_ms, e2e_s, float(b[0].n), float(mapped)
        total_bad, total_checked = state["bad"], n_o

    if rank == 0:
        def finish_line():
            value = reads_per_unit * total_units * args.steps / (dev_ms_max / 1e3)
            e2e_value = reads_per_unit * total_units * args.steps / e2e_max
            out = {
                "metric": metric, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int16/u64", "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": sum(x.h2d_bytes for x in b),
                        "d2h_bytes_per_step": res.d2h_bytes(), "ms_per_step": 1e3 * e2e_max / args.steps,
                        "through": "abg_map_batch (C ABI): host buffers in, host buffers out; the reference arm it is divided "
                                   "by also parses FASTQ and writes SAM -- see fastq_to_sam for the like-for-like number"},
                "gpu_launches": launches * world, "clocks": clocks,
                "%s_per_s" % unit: value / reads_per_unit,
                "%s_mapped_frac" % unit: total_mapped / total_units,
                "parity": {"%s_checked" % unit: total_checked, "ranks_checked": world, "mismatching_records": total_bad,
                           "against": "CPU oracle (oracle/abismal_oracle.cpp), first %d %s of every rank's timed batch" % (n_o, unit)},
            }
            # the seeding is four kernels when binned (hash -> scatter -> filter -> seed_kernel on the survivors),
            # one otherwise; then enum_kernel, dp_kernel, align_kernel and the redo kernel
            times = list(zip(m.SEED_KERNELS, state["seed_ms"])) if m.binned else [("seed_kernel", phase_ms[0])]
            times += list(zip(m.KERNELS[1:], phase_ms[1:]))
            out["kernels"] = {nm: {"ms_per_launch": t / args.steps, "share_of_step": t / dev_ms} for nm, t in times}
            if state["bin_stats"]:
                out["binned_seeding"] = state["bin_stats"]
            # ---- roofline + CPU baseline + front end (rank 0, N = 1) -------------------
            if world == 1 and not args.no_cpu_baseline:
                peak, peak_src = hbm_peak()
                seed_bytes, dp_bytes = algorithmic_bytes(state["oracle_counters"], n_o)
                # dominant kernel: seed_kernel (seed hashing, counter/index lookups, packed compare, candidate sets);
                # its launch time comes from CUDA events recorded on the launching stream between the kernels
                ms_per_launch = phase_ms[0] / args.steps
                achieved = seed_bytes * b[0].n / (ms_per_launch / 1e3) / 1e9
                traffic, gather, traffic_note = None, None, None
                tj = None
                try:
                    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                        tj = json.load(f)
                    # per mode: the ncu --set full capture of the seeding kernels that run (binned: four phases;
                    # ABISMAL_B200_BINS=0: round 1's seed_kernel, kept under "round1_seed_kernel")
                    tj = tj.get(args.mode) if m.binned else tj.get("round1_seed_kernel" if args.mode == "pbat" else None)
                except Exception:
                    tj = None
                if tj is not None:
                    per = float(tj.get("units_per_pair", 1.0))
                    traffic = float(tj["dram_bytes_per_pair"]) * b[0].n / per
                    traffic_note = "%s, %d %s in the capture%s" % (
                        tj.get("source", "profiles/traffic.json"), int(tj["pairs_in_capture"]), unit,
                        "" if int(tj["pairs_in_capture"]) == b[0].n // int(per) else " (scaled to this batch)")
                if tj is not None and not m.binned and "l1_miss_sectors_per_pair" in tj:
                    # one warp per strand: every examined candidate is a random sector, compared with the
                    # random-gather ceiling of the microbenchmark
                    sect_s = float(tj["l1_miss_sectors_per_pair"]) * b[0].n / per / (ms_per_launch / 1e3) / 1e9
                    gather = {"achieved_gsectors_per_s": sect_s,
                              "ceiling_gsectors_per_s": float(tj["random_gather_ceiling_gsectors_per_s"]),
                              "frac": sect_s / float(tj["random_gather_ceiling_gsectors_per_s"]),
                              "note": "32-byte L1-miss sectors per second vs the random-gather ceiling measured with "
                                      "tools/micro/gather_bench3.cu; this, not streaming bandwidth, bounds the kernel"}
                    if "dram_sectors_read_per_pair" in tj and "bucket_contiguous_ceiling_gsectors_per_s" in tj:
                        dsec = float(tj["dram_sectors_read_per_pair"]) * b[0].n / per / (ms_per_launch / 1e3) / 1e9
                        gather["dram_gsectors_per_s"] = dsec
                        gather["bucket_contiguous_ceiling_gsectors_per_s"] = float(tj["bucket_contiguous_ceiling_gsectors_per_s"])
                        gather["dram_frac_of_contiguous_ceiling"] = dsec / float(tj["bucket_contiguous_ceiling_gsectors_per_s"])
                out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                                   "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                                   "peak_source": peak_src,
                                   "algorithmic_bytes_per_%s" % unit[:-1]: seed_bytes,
                                   "algorithmic_bytes_per_%s_whole_path" % unit[:-1]: seed_bytes + dp_bytes,
                                   "whole_step_frac": (seed_bytes + dp_bytes) * b[0].n / (dev_ms / args.steps / 1e3) / 1e9 / peak,
                                   "kernel": ("seeding = hash_kernel + count/prefix/scatter + filter_kernel + seed_kernel (one launch each "
                                              "per batch; the algorithmic bytes of process_seeds are spread over them)")
                                             if m.binned else "seed_kernel",
                                   "ms_per_launch": ms_per_launch, "kernels": out["kernels"],
                                   "counters_from": "CPU oracle on the first %d %s of the batch" % (n_o, unit),
                                   "random_gather": gather}
                n_s = min(args.cpu_sample_pairs, b[0].n)
                s = [prefix + "_sample_%d.fq" % (e + 1) for e in range(len(fqs))]
                for src, dst in zip(fqs, s):
                    sample_fastq(src, dst, 0, n_s)
                ref_sam = prefix + "_sample_ref.sam"
                secs, load = run_reference_map(paths["index"], s, n_cpu, map_flags, ref_sam)
                out["cpu_baseline"] = {
                    "value": reads_per_unit * n_s / secs, "unit": "reads/s", "cores": n_cpu, "kind": "reference",
                    "sample": "first %d %s of the batch, oracle/_ref/abismal map -t %d %s, FASTQ -> SAM, mapping time = wall - "
                              "index loading (%.1fs)" % (n_s, unit, n_cpu, " ".join(map_flags), load),
                    "port_single_thread_reads_per_s": reads_per_unit * n_o / state["port_s"],
                }
                if not args.no_cli:
                    # the same sample through `abismal-b200 map`: SAM identical to the reference binary's (modulo @PG)?
                    ours_sam = prefix + "_sample_ours.sam"
                    run_cli_map(paths["index"], s, map_flags, ours_sam)  # raises when the front end fails
                    n_rec, n_diff = sam_records_differing(ref_sam, ours_sam)
                    out["parity"]["reference_binary"] = {
                        "%s_checked" % unit: n_s, "sam_records": n_rec, "mismatching_records": n_diff,
                        "against": "SAM of the unmodified reference binary (oracle/_ref/abismal map -t %d) on the same FASTQ "
                                   "sample and index, records compared as a sorted multiset, @PG excluded" % n_cpu}
                    # the whole front end on the whole step batch: FASTQ -> SAM (what the reference arm times)
                    big_sam = prefix + "_cli.sam"
                    best = None
                    for _ in range(2):
                        r = run_cli_map(paths["index"], fqs, map_flags, big_sam)
                        if best is None or r[0] < best[0]:
                            best = r
                    out["fastq_to_sam"] = {"value": reads_per_unit * b[0].n / best[0], "unit": "reads/s",
                                           "value_excluding_index_upload": reads_per_unit * b[0].n / max(best[0] - best[2], 1e-9),
                                           "%s" % unit: b[0].n, "seconds": best[0], "index_upload_seconds": best[2],
                                           "stage_busy_time": best[1],
                                           "through": "abismal-b200 map -v %s -i <index> -o <sam> <fastq...>, files on %s; seconds = "
                                                      "the front end's total mapping time, which includes copying the index to "
                                                      "HBM and deriving the seed-context records (a fixed cost per run that "
                                                      "the reference's loading time corresponds to)"
                                                      % (" ".join(map_flags), os.path.dirname(big_sam))}
                    for f in (ours_sam, big_sam):
                        try:
                            os.remove(f)
                        except OSError:
                            pass
                    if n_diff:
                        raise RuntimeError("%d SAM records differ from the reference binary on %d %s" % (n_diff, n_s, unit))
                try:
                    os.remove(ref_sam)
                except OSError:
                    pass
            print(json.dumps(out))
            sys.stdout.flush()
        sync.run("result line (roofline, cpu_baseline, front end)", finish_line)
    else:
        sync.check()
    m.close()
    ix.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
