"""Throughput on a genome with realistic repeat content: a 100 Mbp synthetic genome whose
repeats are modelled on a mammalian one -- a SINE-like family (300 bp, tens of thousands of diverged copies),
a LINE-like family (truncated 6 kb copies), segmental duplications, microsatellites and a tandem satellite --
about a third of the sequence in all.  Reports the per-kernel times of the CUDA path, the share of pairs that
need the redo kernel, the reference binary's time on a sample, and the SAM parity on that sample.
usage: repeat_perf.py [genome_bases] [pairs] [ref_sample_pairs]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = os.path.join(ROOT, "oracle", "_ref", "abismal")
CLI = os.path.join(ROOT, "abismal_b200", "bin", "abismal-b200")
ACGT = np.frombuffer(b"ACGT", np.uint8)


def mutate(rng, seq, div):
    """Substitutions at rate div plus a few short indels."""
    s = seq.copy()
    m = rng.random(s.size) < div
    s[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
    n_indel = rng.poisson(div * s.size / 10.0)
    for _ in range(n_indel):
        p = int(rng.integers(0, max(1, s.size - 8)))
        k = int(rng.integers(1, 6))
        if rng.random() < 0.5:
            s = np.delete(s, slice(p, p + k))
        else:
            s = np.insert(s, p, ACGT[rng.integers(0, 4, k)])
    return s


def repeat_genome(n_bases, seed=5):
    rng = np.random.default_rng(seed)
    g = ACGT[rng.integers(0, 4, n_bases)]
    scale = n_bases / 1e8
    placed = 0

    def put(seq):
        nonlocal placed
        p = int(rng.integers(0, n_bases - seq.size))
        g[p:p + seq.size] = seq
        placed += seq.size
    sine = ACGT[rng.integers(0, 4, 300)]
    for _ in range(int(40000 * scale)):
        put(mutate(rng, sine, rng.uniform(0.05, 0.2)))
    line = ACGT[rng.integers(0, 4, 6000)]
    for _ in range(int(3000 * scale)):
        cut = int(rng.integers(0, 5500))
        put(mutate(rng, line[cut:], rng.uniform(0.02, 0.15)))
    for _ in range(int(20 * scale)):
        seg = ACGT[rng.integers(0, 4, int(rng.integers(20000, 100000)))]
        for _ in range(int(rng.integers(2, 6))):
            put(mutate(rng, seg, rng.uniform(0.01, 0.03)))
    for _ in range(int(2000 * scale)):
        motif = ACGT[rng.integers(0, 4, int(rng.integers(1, 7)))]
        put(mutate(rng, np.tile(motif, int(rng.integers(20, 200))), 0.01))
    mono = ACGT[rng.integers(0, 4, 171)]
    for _ in range(int(10 * scale)):
        arr = np.concatenate([mutate(rng, mono, rng.uniform(0.02, 0.1)) for _ in range(2000)])
        put(arr)
    return g, placed / n_bases


def write_fasta(g, path, n_chroms=8, width=100):
    per = g.size // n_chroms
    per -= per % width
    with open(path, "wb") as f:
        for c in range(n_chroms):
            f.write(b">chr%d\n" % (c + 1))
            x = g[c * per:(c + 1) * per].reshape(-1, width)
            f.write(np.concatenate([x, np.full((x.shape[0], 1), 10, np.uint8)], axis=1).tobytes())


def main():
    n_bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e8)
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    ref_n = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    d = os.environ.get("ABISMAL_B200_CACHE", "/tmp/abismal_b200_bench") + "/repeat_%d" % n_bases
    os.makedirs(d, exist_ok=True)
    fa, idx = d + "/g.fa", d + "/g.idx"
    out = {"genome_bases": n_bases, "pairs": pairs}
    t = time.time()
    if not os.path.exists(fa):
        g, frac = repeat_genome(n_bases)
        write_fasta(g, fa)
        out["repeat_fraction_placed"] = frac
    print("[rep] genome in %.1fs" % (time.time() - t), flush=True)
    t = time.time()
    if not os.path.exists(idx):
        subprocess.check_call([CLI, "idx", fa, idx])
    print("[rep] index (GPU builder) in %.1fs" % (time.time() - t), flush=True)
    pre = d + "/pe"
    if not os.path.exists(pre + "_1.fq"):
        subprocess.check_call([REF, "sim", "-seed", "21", "-l", "150", "-min-fraglen", "150", "-max-fraglen", "400", "-n", str(pairs),
                               "-m", "0.01", "-b", "0.98", "-o", pre, fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    from abismal_b200 import Index, IndexFile, Mapper, workload
    b1, b2 = workload.load_fastq_fast(pre + "_1.fq"), workload.load_fastq_fast(pre + "_2.fq")
    ixf = IndexFile(idx)
    out["index_entries"] = [int(ixf.index_size), int(ixf.index_size_three)]
    ix = Index(ixf, 0)
    for tasks, ovf, tscale in (("1", "256", "4"), ("1", "4096", "128")):
        os.environ["ABISMAL_B200_TASKS"] = tasks
        os.environ["ABISMAL_B200_OVF_PER_ITEM"] = ovf
        os.environ["ABISMAL_B200_TASK_SCALE"] = tscale
        m = Mapper(ix, mode=1, max_batch=b1.n, max_read_len=160)
        m.upload(b1, b2)
        for _ in range(2):
            m.run()
            m.sync()
        res = m.download(b1.n)
        tasks = "%s,ovf=%s,scale=%s" % (tasks, ovf, tscale)
        out["tasks=" + tasks] = {"ms": m.last_kernel_ms, "reads_per_s": 2.0 * b1.n / (m.last_kernel_ms / 1e3),
                                 "kernels_ms": dict(zip(m.KERNELS, m.last_kernel_times)),
                                 "seeding_ms": dict(zip(m.SEED_KERNELS, m.last_seed_times)) if m.binned else None,
                                 "binned_seeding": m.bin_stats() if m.binned else None,
                                 "pairs_mapped_frac": float((res.pe_r1["pos"] != 0).mean()),
                                 "run_stats": m.last_run_stats()}
        print("[rep] tasks=%s %s" % (tasks, json.dumps(out["tasks=" + tasks])), flush=True)
        m.close()
    ix.close()
    for k in ("ABISMAL_B200_TASKS", "ABISMAL_B200_OVF_PER_ITEM", "ABISMAL_B200_TASK_SCALE"):
        os.environ.pop(k, None)  # the front end below runs with its defaults
    # reference binary on a sample + SAM parity through the front end
    s = []
    for e in (1, 2):
        dst = "%s_s_%d.fq" % (pre, e)
        with open("%s_%d.fq" % (pre, e), "rb") as fi, open(dst, "wb") as fo:
            for k, ln in enumerate(fi):
                if k >= 4 * ref_n:
                    break
                fo.write(ln)
        s.append(dst)
    n_cpu = os.cpu_count() or 1
    t = time.perf_counter()
    subprocess.check_call([REF, "map", "-t", str(n_cpu), "-i", idx, "-o", d + "/ref.sam"] + s, stderr=subprocess.DEVNULL)
    ref_s = time.perf_counter() - t
    t = time.perf_counter()
    p = subprocess.run([CLI, "map", "-i", idx, "-o", d + "/ours.sam"] + s, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        raise RuntimeError("abismal-b200 map failed: " + p.stderr[-1500:])
    ours_s = time.perf_counter() - t
    a = sorted(ln for ln in open(d + "/ref.sam", "rb") if not ln.startswith(b"@PG"))
    b = sorted(ln for ln in open(d + "/ours.sam", "rb") if not ln.startswith(b"@PG"))
    out["reference"] = {"pairs": ref_n, "cores": n_cpu, "wall_s": ref_s, "reads_per_s_wall": 2.0 * ref_n / ref_s,
                        "front_end_wall_s": ours_s, "sam_identical": a == b, "sam_records": len(a)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
