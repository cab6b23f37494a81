"""Randomised pinning of the oracle (CPU only): random repeat genomes (with and without IUPAC codes), reads from
the reference's `sim` with random length / error rate / conversion / fragment range, a random set of map flags;
the SAM and the statistics of oracle/oracle_map (the restatement behind the product's host code) must equal the
unmodified reference binary's.  usage: fuzz_oracle.py [iterations] [first_seed] [oracle|cli]
(`cli`: the product's `abismal-b200 map` instead -- needs a GPU)"""
import os, random, shutil, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers, make_genome

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
helpers.ensure_built()
TOOL = helpers.CLI if (len(sys.argv) > 3 and sys.argv[3] == "cli") else helpers.ORACLE_MAP
bad = 0
for it in range(n_iter):
    rnd = random.Random(seed0 + it)
    d = tempfile.mkdtemp(prefix="fuzz_", dir="/tmp")
    try:
        ws = helpers.Workspace(d)
        iupac = rnd.random() < 0.5
        make_genome.write_fasta(make_genome.repeat_genome(seed=seed0 + it, scale=rnd.choice([0.05, 0.1, 0.2]), iupac=iupac), ws.path("g.fa"))
        ws.ref("idx", "tests/g.fa", "tests/g.idx")
        length = rnd.choice([50, 64, 75, 100, 150, 151, 200, 250])
        paired = rnd.random() < 0.6
        conv = rnd.choice([None, "-a", "-R"])
        cmd = ["sim", "-seed", str(rnd.randrange(1, 10 ** 6)), "-l", str(length), "-n", str(rnd.choice([500, 1500])),
               "-m", rnd.choice(["0", "0.01", "0.03", "0.06"]), "-b", rnd.choice(["0.5", "0.98", "1.0"]), "-o", "tests/r"]
        if paired:
            lo = rnd.choice([length, length + 20])
            cmd += ["-min-fraglen", str(lo), "-max-fraglen", str(lo + rnd.choice([50, 300]))]
        else:
            cmd.append("-single")
        if conv:
            cmd.append(conv)
        ws.ref(*(cmd + ["tests/g.fa"]))
        flags = []
        if conv == "-a":
            flags.append("-P" if paired else "-A")
        if conv == "-R":
            flags.append("-R")
        if rnd.random() < 0.3:
            flags.append("-a")
        if rnd.random() < 0.3:
            flags += ["-m", rnd.choice(["0.05", "0.2"])]
        if rnd.random() < 0.3:
            flags += ["-c", rnd.choice(["3", "20", "500"])]
        if paired and rnd.random() < 0.3:
            flags += ["-l", "40", "-L", rnd.choice(["250", "1000"])]
        files = ["tests/r_1.fq", "tests/r_2.fq"] if paired else ["tests/r_1.fq"]
        args = flags + ["-i", "tests/g.idx"] + files
        rs, rt, _ = ws.map_with(helpers.REF_BIN, "ref", args)
        os_, ot, _ = ws.map_with(TOOL, "or", args)
        same = helpers.sam_body(rs) == helpers.sam_body(os_) and open(rt).read() == open(ot).read()
        print("[fuzz %d] iupac=%d len=%d paired=%d conv=%s flags=%s records=%d -> %s"
              % (seed0 + it, iupac, length, paired, conv, " ".join(flags), len(helpers.sam_body(rs)), "ok" if same else "DIFFER"), flush=True)
        if not same:
            bad += 1
            shutil.copytree(d, "/tmp/fuzz_fail_%d" % (seed0 + it), dirs_exist_ok=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
print("fuzz: %d of %d differ" % (bad, n_iter))
sys.exit(1 if bad else 0)
