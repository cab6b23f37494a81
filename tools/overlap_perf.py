"""Launch-shape timing on the bench workload: sequential vs overlapped seed/align launch, CTAs per SM.
usage: overlap_perf.py [pairs] [configs]   config = OVERLAP[:SEED_BLOCKS[:AUX_BLOCKS[:MINB[:MINB_SEED[:CHUNK]]]]], comma separated
ABISMAL_B200_LIB selects a differently compiled library (one per process)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: F401
from abismal_b200 import workload, Index, Mapper, MODE_A_RICH, MODE_PAIRED
from abismal_b200.capi import Results
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
configs = (sys.argv[2] if len(sys.argv) > 2 else "0,1").split(",")
check_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tag = os.path.basename(os.environ.get("ABISMAL_B200_LIB", "default"))
log = lambda *a: print("[op %s]" % tag, *a, flush=True)
ixf, paths = workload.get_index(int(3.1e9), 20251017, device=0, need_files=True, log=log)
ref_bin = os.path.join(ROOT, "oracle", "_ref", "abismal")
prefix = os.path.join(paths["dir"], "pbat_n%d_r0" % pairs)
fq1, fq2 = workload.simulate_reads(ref_bin, paths["fasta"], prefix, pairs, seed=20251017 % 1000, paired=True,
                                   mode_flag="-a", n_procs=16, log=log)
b1, b2 = workload.load_fastq_fast(fq1), workload.load_fastq_fast(fq2)
mode = MODE_PAIRED | MODE_A_RICH
ix = Index(ixf, 0)
first = None
for cfg in configs:
    f = cfg.split(":")
    os.environ["ABISMAL_B200_OVERLAP"] = f[0]
    for k, name in ((1, "ABISMAL_B200_SEED_BLOCKS"), (2, "ABISMAL_B200_AUX_BLOCKS"), (3, "ABISMAL_B200_MINB"),
                    (4, "ABISMAL_B200_MINB_SEED"), (5, "ABISMAL_B200_CHUNK")):
        if len(f) > k and f[k] != "":
            os.environ[name] = f[k]
        else:
            os.environ.pop(name, None)
    m = Mapper(ix, mode=mode, max_batch=b1.n, max_read_len=max(b1.max_len, b2.max_len, 64))
    m.upload(b1, b2); m.sync()
    ms, ph = [], None
    for it in range(4):
        m.run(); m.sync(); ms.append(m.last_kernel_ms); ph = m.last_phase_ms
    res = Results(b1.n, True, m.stride, pinned=True)
    p1, p2 = b1.to_pinned(), b2.to_pinned()
    m.map_batch(p1, p2, res)
    t = time.perf_counter()
    for _ in range(3):
        m.map_batch(p1, p2, res)
    e2e = (time.perf_counter() - t) / 3
    log("config %s: kernel ms %s (phases %s) -> %.3f M pairs/s kernel, e2e %.1f ms -> %.3f M pairs/s"
        % (cfg, ["%.1f" % x for x in ms], ["%.1f" % x for x in ph], b1.n / min(ms[1:]) / 1e3, e2e * 1e3, b1.n / e2e / 1e6))
    if first is None:
        first = res
    else:
        import helpers
        try:
            helpers.assert_results_equal(res, first, True)
            log("config %s results identical to config %s" % (cfg, configs[0]))
        except AssertionError as e:
            log("MISMATCH: config %s differs from %s: %s" % (cfg, configs[0], str(e)[:400]))
    m.close()
if check_n:
    import helpers
    o = helpers.OracleMapper(ixf, mode=mode)
    want = o.map_batch(b1.slice(0, check_n), b2.slice(0, check_n))
    m = Mapper(ix, mode=mode, max_batch=check_n, max_read_len=max(b1.max_len, b2.max_len, 64))
    got = m.map_batch(b1.slice(0, check_n), b2.slice(0, check_n))
    helpers.assert_results_equal(got, want, True)
    log("parity vs oracle ok on %d pairs" % check_n)
