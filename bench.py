#!/usr/bin/env python
"""bench.py -- throughput of the `abismal map` hot path on B200.

Workload (BASELINE.json configs[3] shape): synthetic i.i.d. 3.1 Gbp genome,
150 bp paired-end PBAT reads from the reference's `sim -a`, mapped with -P.
A "step" maps one batch of --pairs read pairs per GPU; reads are sharded over
GPUs with the index replicated (weak scaling, no data-path collective; NCCL
only sums the mapping statistics).

  value  reads/s with the batch already resident in HBM (CUDA-event time of
         the K launches, max over ranks)
  e2e    reads/s through the public call (abg_map_batch: host buffers in,
         host buffers out, H2D/D2H inside the timed region)
  --impl reference   the unmodified reference binary (oracle/_ref/abismal map
         -t <all cores>) on a bounded sample of the same reads and index.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mapped reads/sec (150bp PE bisulfite, 3.1 Gbp genome)"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "abismal")


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


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


def sample_fastq(src, dst, n_records):
    with open(src, "rb") as fi, open(dst, "wb") as fo:
        for k, ln in enumerate(fi):
            if k >= 4 * n_records:
                break
            fo.write(ln)


def run_reference_map(index_path, fq1, fq2, n_threads, extra=("-P",)):
    """-> (mapping seconds, loading seconds); mapping = wall - index loading (-v log line)."""
    out = os.path.join(os.path.dirname(fq1), "ref_sample.sam")
    cmd = [REF_BIN, "map", "-v", "-t", str(n_threads)] + list(extra) + ["-i", index_path, "-o", out, fq1] + ([fq2] if fq2 else [])
    t = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t
    if p.returncode != 0:
        raise RuntimeError("reference map failed: " + p.stderr[-500:])
    load = 0.0
    for ln in p.stderr.splitlines():
        if "loading time:" in ln:
            load = float(ln.split("loading time:")[1].strip().rstrip("s"))
    try:
        os.remove(out)
    except OSError:
        pass
    return wall - load, load


def algorithmic_bytes(counters, units):
    """bytes = 8 N_lookup + 4 N_entry + 8 (N_word + N_cmp) + 0.5 N_dpref, per unit.
    -> (seeding part: counter probes, index entries, packed compares; alignment part: DP reference bases)"""
    c = counters
    return ((8 * c["n_lookup"] + 4 * c["n_entry"] + 8 * (c["n_word"] + c["n_cmp"])) / units,
            0.5 * c["n_dpref"] / units)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-bases", type=float, default=3.1e9)
    ap.add_argument("--pairs", type=int, default=1 << 20, help="read pairs per GPU per step")
    ap.add_argument("--cpu-sample-pairs", type=int, default=200000)
    ap.add_argument("--seed", type=int, default=20251017)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    genome_bases = int(args.genome_bases)
    n_cpu = os.cpu_count() or 1

    if args.impl == "reference" and rank != 0:
        return 0

    import torch  # plumbing: device selection, NCCL, barriers
    import torch.distributed as dist
    from abismal_b200 import workload

    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device to prepare the synthetic index"}))
            return 0
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload (untimed): genome, index, reads --------------------------------
    paths = None
    ixf = None
    if rank == 0:
        ixf, paths = workload.get_index(genome_bases, args.seed, device=local_rank, need_files=True, log=log)
    if use_dist:
        dist.barrier()
    if ixf is None:
        ixf, paths = workload.get_index(genome_bases, args.seed, device=local_rank, need_files=True, log=log)
    sim_procs = max(1, min(16, n_cpu // max(world, 1)))
    prefix = os.path.join(paths["dir"], "pbat_n%d_r%d" % (args.pairs, rank))
    fq1, fq2 = workload.simulate_reads(REF_BIN, paths["fasta"], prefix, args.pairs, seed=args.seed % 1000 + rank,
                                       paired=True, mode_flag="-a", n_procs=sim_procs, log=log)

    config = {
        "workload": "synthetic i.i.d. %.2f Gbp genome (24 chroms), 150bp PE PBAT reads (sim -a -m 0.01 -b 0.98, "
                    "fragments 150-400), abismal map -P; one batch of %d pairs per GPU per step "
                    "(BASELINE configs[3] shape)" % (genome_bases / 1e9, args.pairs),
        "pairs_per_gpu_per_step": args.pairs,
        "index": "replicated per GPU, built on GPU (byte-identical to `abismal idx`)",
        "parallelism": "reads sharded x%d, no data-path collective" % world,
        "l2": "inputs larger than L2: %.1f GB index gathered at random + %.0f MB of reads per step"
              % (2.7 * genome_bases / 3.1e9, args.pairs * 300 / 1e6),
    }

    if args.impl == "reference":
        # ---- reference arm: unmodified reference binary, all host cores, bounded sample ----
        n_s = min(args.cpu_sample_pairs, args.pairs)
        s1, s2 = prefix + "_sample_1.fq", prefix + "_sample_2.fq"
        sample_fastq(fq1, s1, n_s)
        sample_fastq(fq2, s2, n_s)
        for _ in range(args.warmup):
            run_reference_map(paths["index"], s1, s2, n_cpu)
        secs = [run_reference_map(paths["index"], s1, s2, n_cpu)[0] for _ in range(args.steps)]
        total = sum(secs)
        value = 2.0 * n_s * args.steps / total
        sample = "first %d pairs of the step batch per step, abismal map -t %d -P (mapping time = wall - index loading)" % (n_s, n_cpu)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/u64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": n_cpu, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ---- our arm -----------------------------------------------------------------
    from abismal_b200 import Index, Mapper, MODE_A_RICH, MODE_PAIRED
    from abismal_b200.capi import Results
    t = time.time()
    b1, b2 = workload.load_fastq_fast(fq1), workload.load_fastq_fast(fq2)
    log("reads loaded in %.1fs" % (time.time() - t))
    t = time.time()
    ix = Index(ixf, local_rank)
    log("index resident in HBM: %.2f GB, uploaded in %.1fs" % (ix.device_bytes / 1e9, time.time() - t))
    mode = MODE_PAIRED | MODE_A_RICH
    m = Mapper(ix, mode=mode, max_batch=b1.n, max_read_len=max(b1.max_len, b2.max_len, 64))
    res = Results(b1.n, True, m.stride, pinned=True)
    b1, b2 = b1.to_pinned(), b2.to_pinned()  # the e2e leg copies each step's inputs from pinned host memory

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # device-resident leg
    m.upload(b1, b2)
    m.sync()
    for _ in range(args.warmup):
        m.run()
        m.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    kernel_ms, phase_ms = [], [0.0, 0.0, 0.0]
    for _ in range(args.steps):
        m.run()
        m.sync()
        kernel_ms.append(m.last_kernel_ms)
        phase_ms = [a + b for a, b in zip(phase_ms, m.last_phase_ms)]
    barrier()
    dev_ms = sum(kernel_ms)
    launches = args.steps * m.launches_per_run

    # end-to-end leg through the public call
    for _ in range(2):
        m.map_batch(b1, b2, res)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.map_batch(b1, b2, res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()

    mapped_pairs = int((res.pe_r1["pos"] != 0).sum())
    stats = torch.tensor([dev_ms, e2e_s, float(b1.n), float(mapped_pairs)], dtype=torch.float64, device="cuda")
    if use_dist:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)  # NCCL: the only collective (mapping statistics)
        dev_ms_max, e2e_max = float(mx[0]), float(mx[1])
        total_pairs, total_mapped = float(sm[2]), float(sm[3])
    else:
        dev_ms_max, e2e_max, total_pairs, total_mapped = dev_ms, e2e_s, float(b1.n), float(mapped_pairs)

    if rank == 0:
        value = 2.0 * total_pairs * args.steps / (dev_ms_max / 1e3)
        e2e_value = 2.0 * total_pairs * args.steps / e2e_max
        out = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16/u64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": b1.h2d_bytes + b2.h2d_bytes,
                    "d2h_bytes_per_step": res.d2h_bytes(), "ms_per_step": 1e3 * e2e_max / args.steps},
            "gpu_launches": launches * world, "clocks": clocks,
            "pairs_mapped_frac": total_mapped / total_pairs,
        }
        config["pairs_per_s"] = value / 2.0
        # ---- roofline + CPU baseline (rank 0, N = 1) -----------------------------
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import helpers
            peak, peak_src = hbm_peak()
            n_o = min(4000, b1.n)
            o = helpers.OracleMapper(ixf, mode=mode)
            t = time.perf_counter()
            want = o.map_batch(b1.slice(0, n_o), b2.slice(0, n_o))
            port_s = time.perf_counter() - t
            # parity at the full configuration: the records the timed e2e leg produced for these pairs against
            # the CPU oracle (checker only; position, flags, NM of pe.r1/pe.r2/se1/se2 and the CIGAR lengths)
            bad = 0
            for k in ("se1", "n_cigar1", "pe_r1", "pe_r2", "se2", "n_cigar2"):
                bad += int((getattr(res, k)[:n_o] != getattr(want, k)[:n_o]).sum())
            out["parity"] = {"pairs_checked": n_o, "mismatching_records": bad, "against": "CPU oracle (oracle/abismal_oracle.cpp)"}
            if bad:
                raise SystemExit("bench: %d result records differ from the oracle on the first %d pairs" % (bad, n_o))
            seed_bytes, dp_bytes = algorithmic_bytes(o.counters.as_dict(), n_o)
            o.close()
            # dominant kernel: seed_kernel (seed hashing, counter/index lookups, packed compare, candidate sets);
            # its launch time comes from CUDA events recorded on the launching stream between the kernels
            names = ("seed_kernel", "align_kernel", "map_reads_kernel(redo)")
            per_kernel = {nm: {"ms_per_launch": t / args.steps, "share_of_step": t / dev_ms}
                          for nm, t in zip(names, phase_ms)}
            ms_per_launch = phase_ms[0] / args.steps
            bytes_per_pair = seed_bytes
            achieved = bytes_per_pair * b1.n / (ms_per_launch / 1e3) / 1e9
            traffic, gather = None, None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    tj = json.load(f)
                if not tj.get("kernel", "").startswith("seed_kernel"):
                    raise KeyError("traffic.json is not a seed_kernel capture")
                # ncu --set full capture of the same kernel on a 262144-pair batch of the same reads,
                # scaled to this launch's batch (the kernel's work is linear in the number of pairs)
                traffic = float(tj["dram_bytes_per_pair"]) * b1.n
                sect_s = float(tj["l1_miss_sectors_per_pair"]) * b1.n / (ms_per_launch / 1e3) / 1e9
                gather = {"achieved_gsectors_per_s": sect_s,
                          "ceiling_gsectors_per_s": float(tj["random_gather_ceiling_gsectors_per_s"]),
                          "frac": sect_s / float(tj["random_gather_ceiling_gsectors_per_s"]),
                          "note": "32-byte L1-miss sectors per second vs the random-gather ceiling measured with "
                                  "tools/micro/gather_bench3.cu; this, not streaming bandwidth, bounds the kernel"}
                if "dram_sectors_read_per_pair" in tj and "bucket_contiguous_ceiling_gsectors_per_s" in tj:
                    dsec = float(tj["dram_sectors_read_per_pair"]) * b1.n / (ms_per_launch / 1e3) / 1e9
                    gather["dram_gsectors_per_s"] = dsec
                    gather["bucket_contiguous_ceiling_gsectors_per_s"] = float(tj["bucket_contiguous_ceiling_gsectors_per_s"])
                    gather["dram_frac_of_contiguous_ceiling"] = dsec / float(tj["bucket_contiguous_ceiling_gsectors_per_s"])
            except Exception:
                pass
            out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                               "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                               "algorithmic_bytes_per_pair": bytes_per_pair,
                               "algorithmic_bytes_per_pair_whole_path": seed_bytes + dp_bytes,
                               "kernel": "seed_kernel", "ms_per_launch": ms_per_launch, "kernels": per_kernel,
                               "counters_from": "CPU oracle on the first %d pairs of the batch" % n_o,
                               "random_gather": gather}
            n_s = min(args.cpu_sample_pairs, b1.n)
            s1, s2 = prefix + "_sample_1.fq", prefix + "_sample_2.fq"
            sample_fastq(fq1, s1, n_s)
            sample_fastq(fq2, s2, n_s)
            secs, load = run_reference_map(paths["index"], s1, s2, n_cpu)
            out["cpu_baseline"] = {
                "value": 2.0 * n_s / secs, "unit": "reads/s", "cores": n_cpu, "kind": "reference",
                "sample": "first %d pairs of the batch, oracle/_ref/abismal map -t %d -P, mapping time = wall - "
                          "index loading (%.1fs)" % (n_s, n_cpu, load),
                "port_single_thread_reads_per_s": 2.0 * n_o / port_s,
            }
        print(json.dumps(out))
    m.close()
    ix.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
