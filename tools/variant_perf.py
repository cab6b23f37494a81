"""Kernel-variant timing on the bench workload (3.1 Gbp synthetic genome, PBAT pairs).
usage: variant_perf.py [pairs] [variants, e.g. 2,3,4 or 3/0/0,3/1/0,3/1/1] [check_n]
A variant is MINB[n][/CTX[/CC]]: n = single-kernel path, CTX / CC = ABISMAL_B200_CTX / ABISMAL_B200_CC
(seed-context records, compact counters) for the index the variant runs on."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from abismal_b200 import workload, Index, Mapper, MODE_A_RICH, MODE_PAIRED
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
variants = (sys.argv[2] if len(sys.argv) > 2 else "3,3n").split(",")  # MINB, suffix n = single-kernel path
check_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
log = lambda *a: print("[vp]", *a, flush=True)
ixf, paths = workload.get_index(int(3.1e9), 20251017, device=0, need_files=True, log=log)
ref_bin = os.path.join(ROOT, "oracle", "_ref", "abismal")
prefix = os.path.join(paths["dir"], "pbat_n%d_r0" % pairs)
fq1, fq2 = workload.simulate_reads(ref_bin, paths["fasta"], prefix, pairs, seed=20251017 % 1000, paired=True,
                                   mode_flag="-a", n_procs=16, log=log)
b1, b2 = workload.load_fastq_fast(fq1), workload.load_fastq_fast(fq2)
mode = MODE_PAIRED | MODE_A_RICH
first = None
ix, ix_key = None, None
for vfull in variants:
    parts = vfull.split("/")
    v = parts[0]
    key = (parts[1] if len(parts) > 1 else "1", parts[2] if len(parts) > 2 else "1")
    if key != ix_key:
        if ix is not None:
            ix.close()
        os.environ["ABISMAL_B200_CTX"], os.environ["ABISMAL_B200_CC"] = key
        t = time.time()
        ix = Index(ixf, 0)
        ix_key = key
        log("index ctx=%s cc=%s: %.2f GB resident, created in %.1fs" % (key[0], key[1], ix.device_bytes / 1e9, time.time() - t))
    os.environ["ABISMAL_B200_MINB"] = v.rstrip("n")
    os.environ["ABISMAL_B200_SPLIT"] = "0" if v.endswith("n") else "1"
    m = Mapper(ix, mode=mode, max_batch=b1.n, max_read_len=max(b1.max_len, b2.max_len, 64),
               count_work=bool(os.environ.get("COUNT")))
    m.upload(b1, b2); m.sync()
    ms = []
    for it in range(4):
        m.run(); m.sync(); ms.append(m.last_kernel_ms)
    log("variant %s kernel ms %s -> %.3f M pairs/s" % (vfull, ["%.1f" % x for x in ms], b1.n / min(ms[1:]) / 1e3))
    if os.environ.get("COUNT"):
        log(m.counters().as_dict())
    res = m.map_batch(b1, b2)
    if first is None:
        first = res
    else:
        import helpers
        try:
            helpers.assert_results_equal(res, first, True)
            log("variant %s results identical to variant %s" % (vfull, variants[0]))
        except AssertionError as e:
            log("MISMATCH: variant %s differs from variant %s: %s" % (vfull, variants[0], str(e)[:500]))
    m.close()
if check_n:
    import helpers
    o = helpers.OracleMapper(ixf, mode=mode)
    t = time.time(); want = o.map_batch(b1.slice(0, check_n), b2.slice(0, check_n)); dt = time.time() - t
    v0 = variants[-1].split("/")[0]  # the index of the last variant is still resident
    os.environ["ABISMAL_B200_MINB"] = v0.rstrip("n")
    os.environ["ABISMAL_B200_SPLIT"] = "0" if v0.endswith("n") else "1"
    m = Mapper(ix, mode=mode, max_batch=check_n, max_read_len=max(b1.max_len, b2.max_len, 64))
    got = m.map_batch(b1.slice(0, check_n), b2.slice(0, check_n))
    helpers.assert_results_equal(got, want, True)
    log("parity vs oracle ok on %d pairs (oracle %.0f pairs/s single thread)" % (check_n, check_n / dt))
