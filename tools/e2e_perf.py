"""End-to-end (abg_map_batch) timing vs sub-batch size and buffer kind on the bench workload.
usage: e2e_perf.py [pairs] [chunks csv]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from abismal_b200 import workload, Index, Mapper, MODE_A_RICH, MODE_PAIRED
from abismal_b200.capi import Results
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
chunks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,65536,16384").split(",")]
log = lambda *a: print("[e2e]", *a, flush=True)
ixf, paths = workload.get_index(int(3.1e9), 20251017, device=0, need_files=True, log=log)
ref_bin = os.path.join(ROOT, "oracle", "_ref", "abismal")
prefix = os.path.join(paths["dir"], "pbat_n%d_r0" % pairs)
fq1, fq2 = workload.simulate_reads(ref_bin, paths["fasta"], prefix, pairs, seed=20251017 % 1000, paired=True,
                                   mode_flag="-a", n_procs=16, log=log)
b1, b2 = workload.load_fastq_fast(fq1), workload.load_fastq_fast(fq2)
p1, p2 = b1.to_pinned(), b2.to_pinned()
ix = Index(ixf, 0)
mode = MODE_PAIRED | MODE_A_RICH
for c in chunks:
    os.environ["ABISMAL_B200_CHUNK"] = str(c if c else pairs)
    m = Mapper(ix, mode=mode, max_batch=b1.n, max_read_len=max(b1.max_len, b2.max_len, 64))
    m.upload(b1, b2); m.sync(); m.run(); m.sync(); m.run(); m.sync()
    kms = m.last_kernel_ms
    for name, (x1, x2), pin in (("pageable", (b1, b2), False), ("pinned", (p1, p2), True)):
        res = Results(b1.n, True, m.stride, pinned=pin)
        m.map_batch(x1, x2, res)
        ts = []
        for _ in range(3):
            t = time.perf_counter(); m.map_batch(x1, x2, res); ts.append((time.perf_counter() - t) * 1e3)
        log("chunk %8d %-8s map_batch ms %s  (device-resident kernel %.1f ms)" % (c if c else pairs, name, ["%.1f" % t for t in ts], kms))
    m.close()
