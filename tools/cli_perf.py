"""FASTQ -> SAM/BAM wall time of the `abismal-b200 map` front end on the bench workload (3.1 Gbp synthetic
genome, PBAT pairs), next to the reference binary on a sample of the same files.
usage: cli_perf.py [pairs_per_file_copy] [copies] [ref_sample_pairs]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abismal_b200 import workload
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ref_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
log = lambda *a: print("[cli]", *a, flush=True)
ixf, paths = workload.get_index(int(3.1e9), 20251017, device=0, need_files=True, log=log)
del ixf
ref_bin = os.path.join(ROOT, "oracle", "_ref", "abismal")
cli = os.path.join(ROOT, "abismal_b200", "bin", "abismal-b200")
prefix = os.path.join(paths["dir"], "pbat_n%d_r0" % pairs)
fq1, fq2 = workload.simulate_reads(ref_bin, paths["fasta"], prefix, pairs, seed=20251017 % 1000, paired=True,
                                   mode_flag="-a", n_procs=16, log=log)
big = []
for k, fq in enumerate((fq1, fq2)):
    dst = "/dev/shm/cli_perf_%d.fq" % (k + 1)
    with open(dst, "wb") as fo:
        data = open(fq, "rb").read()
        for _ in range(copies):
            fo.write(data)
    big.append(dst)
n_pairs = pairs * copies
out = {}
def run(tag, cmd, n):
    t = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t
    if p.returncode != 0:
        log(tag, "FAILED", p.stderr[-1500:]); return
    info = {}
    for ln in p.stderr.splitlines():
        for key in ("loading time:", "total mapping time:"):
            if key in ln:
                info[key.rstrip(":")] = float(ln.split(key)[1].strip().rstrip("s"))
        if "index upload to HBM:" in ln:
            info["index upload"] = float(ln.split("index upload to HBM:")[1].split("s")[0])
        if "engine set-up" in ln:
            info["engine set-up"] = float(ln.split("):")[1].strip().rstrip("s"))
        if "stage busy time" in ln:
            info["stages"] = ln.split("stage busy time:")[1].strip()
    mt = info.get("total mapping time", wall - info.get("loading time", 0.0))
    out[tag] = {"wall_s": wall, "pairs": n, "reads_per_s_mapping": 2 * n / mt, "reads_per_s_wall": 2 * n / wall, **info}
    log(tag, json.dumps(out[tag]))
for tag, extra, dst in (("sam_t16", ["-t", "16"], "/dev/shm/cli_perf.sam"), ("sam_t16_again", ["-t", "16"], "/dev/shm/cli_perf.sam"),
                        ("sam_t4", ["-t", "4"], "/dev/shm/cli_perf.sam"), ("sam_devnull", ["-t", "16"], "/dev/null"),
                        ("bam_t16", ["-B", "-t", "16"], "/dev/shm/cli_perf.bam")):
    run(tag, [cli, "map", "-v", "-P"] + extra + ["-i", paths["index"], "-o", dst, "-s", "/dev/shm/cli_perf.stats"] + big, n_pairs)
if ref_n:
    s = []
    for k, fq in enumerate(big):
        dst = "/dev/shm/cli_ref_%d.fq" % (k + 1)
        with open(fq, "rb") as fi, open(dst, "wb") as fo:
            for i, ln in enumerate(fi):
                if i >= 4 * ref_n: break
                fo.write(ln)
        s.append(dst)
    run("reference_t%d" % os.cpu_count(), [ref_bin, "map", "-v", "-P", "-t", str(os.cpu_count()), "-i", paths["index"],
                                           "-o", "/dev/shm/cli_ref.sam", "-s", "/dev/shm/cli_ref.stats"] + s, ref_n)
    # same sample through the GPU front end: SAM must be identical apart from @PG
    run("ours_on_ref_sample", [cli, "map", "-v", "-P", "-i", paths["index"], "-o", "/dev/shm/cli_ours.sam",
                               "-s", "/dev/shm/cli_ours.stats"] + s, ref_n)
    # the reference with -t > 1 writes its 1000-read batches in the order threads finish: compare as multisets
    a = sorted(ln for ln in open("/dev/shm/cli_ref.sam") if not ln.startswith("@PG"))
    b = sorted(ln for ln in open("/dev/shm/cli_ours.sam") if not ln.startswith("@PG"))
    out["sam_identical_to_reference"] = a == b
    out["stats_identical_to_reference"] = open("/dev/shm/cli_ref.stats").read() == open("/dev/shm/cli_ours.stats").read()
    log("SAM identical to the reference binary on %d pairs: %s, stats identical: %s" % (ref_n, a == b, out["stats_identical_to_reference"]))
print(json.dumps(out))
for f in os.listdir("/dev/shm"):
    if f.startswith("cli_"):
        os.remove(os.path.join("/dev/shm", f))
