"""Scratch perf probe: random genome of G bases, reference idx + sim, kernel timing."""
import os, sys, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers, make_genome
from abismal_b200 import Index, IndexFile, Mapper, load_fastq

G = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
d = tempfile.mkdtemp(dir="/tmp")
ws = helpers.Workspace(d)
t = time.time()
rng = np.random.default_rng(5)
with open(ws.path("g.fa"), "w") as f:
    for c in range(4):
        f.write(">chr%d\n" % (c + 1))
        s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, G // 4)]
        s = s.reshape(-1, 100) if (G // 4) % 100 == 0 else s
        if s.ndim == 2:
            out = np.concatenate([s, np.full((s.shape[0], 1), 10, np.uint8)], axis=1)
            f.write(out.tobytes().decode())
        else:
            f.write(s.tobytes().decode() + "\n")
print("genome %.1fs" % (time.time() - t)); t = time.time()
for nt in (os.cpu_count(), 8, 4, 1):  # the reference's block partition crashes for some thread counts
    try:
        ws.ref("idx", "-t", str(nt), "tests/g.fa", "tests/g.idx"); break
    except RuntimeError as e:
        print("idx -t %d failed" % nt)
print("idx %.1fs" % (time.time() - t)); t = time.time()
if mode & 1:
    ws.ref("sim", "-seed", "9", "-l", "150", "-min-fraglen", "150", "-max-fraglen", "400", "-n", str(NP), "-m", "0.01", "-b", "0.98", "-o", "tests/r", "tests/g.fa")
else:
    ws.ref("sim", "-single", "-seed", "9", "-l", "150", "-n", str(NP), "-m", "0.01", "-b", "0.98", "-o", "tests/r", "tests/g.fa")
print("sim %.1fs" % (time.time() - t)); t = time.time()
ixf = IndexFile(ws.path("g.idx"))
b1 = load_fastq(ws.path("r_1.fq")); b2 = load_fastq(ws.path("r_2.fq")) if mode & 1 else None
print("load %.1fs" % (time.time() - t)); t = time.time()
ix = Index(ixf, 0)
print("index on device: %.2f GB, %.1fs" % (ix.device_bytes / 1e9, time.time() - t))
m = Mapper(ix, mode=mode, max_batch=b1.n, max_read_len=160, count_work=bool(int(os.environ.get("COUNT", "0"))))
args = (b1, b2) if mode & 1 else (b1,)
m.upload(*args)
for it in range(4):
    m.run(); m.sync()
    print("kernel %.2f ms -> %.0f reads/s" % (m.last_kernel_ms, (2 if mode & 1 else 1) * b1.n / (m.last_kernel_ms / 1e3)))
t = time.time(); res = m.map_batch(*args); dt = time.time() - t
print("e2e map_batch %.1f ms" % (dt * 1e3))
if os.environ.get("COUNT"):
    print(m.counters())
if os.environ.get("CHECK"):
    o = helpers.OracleMapper(ixf, mode=mode)
    n = int(os.environ["CHECK"])
    t = time.time(); want = o.map_batch(*[x.slice(0, n) for x in args]); dt = time.time() - t
    got = m.map_batch(*[x.slice(0, n) for x in args])
    helpers.assert_results_equal(got, want, bool(mode & 1))
    print("oracle 1 thread: %.0f reads/s; parity ok on %d" % ((2 if mode & 1 else 1) * n / dt, n), o.counters.as_dict())
if os.environ.get("REFT"):
    t = time.time(); ws.ref("map", "-t", str(os.cpu_count()), "-i", "tests/g.idx", "-o", "tests/ref.sam", "tests/r_1.fq", *( ["tests/r_2.fq"] if mode & 1 else []))
    print("reference map -t %d: %.1fs total" % (os.cpu_count(), time.time() - t))
import shutil; shutil.rmtree(d)
