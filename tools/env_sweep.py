"""Environment-variable sweep on the bench workload (3.1 Gbp synthetic genome): one index upload, one set of
reads, one Mapper per variant (the tuning variables are read by abg_mapper_create).
usage: env_sweep.py MODE PAIRS "A=1,B=2;A=3;..." [check_n]      (an empty variant = the defaults)
Prints the per-kernel times of every variant, whether its results equal the first variant's, and (check_n)
parity of the last variant against the CPU oracle on the first check_n items.  Shares bench.py's cache."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: F401  (device context like bench.py)
from abismal_b200 import workload, Index, Mapper, MODE_A_RICH, MODE_PAIRED, MODE_RANDOM_PBAT
mode_name = sys.argv[1] if len(sys.argv) > 1 else "pbat"
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
variants = (sys.argv[3] if len(sys.argv) > 3 else "").split(";")
check_n = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sim_flag, mode, paired = {"pbat": ("-a", MODE_PAIRED | MODE_A_RICH, True), "rpbat": ("-R", MODE_PAIRED | MODE_RANDOM_PBAT, True),
                          "se": (None, 0, False)}[mode_name]
log = lambda *a: print("[sweep]", *a, flush=True)
ixf, paths = workload.get_index(int(3.1e9), 20251017, device=0, need_files=True, log=log)
n_units = pairs if paired else 2 * pairs
prefix = os.path.join(paths["dir"], "%s_n%d_r0" % (mode_name, n_units))
fqs = workload.simulate_reads(os.path.join(ROOT, "oracle", "_ref", "abismal"), paths["fasta"], prefix, n_units, seed=20251017 % 1000,
                              paired=paired, mode_flag=sim_flag, n_procs=min(16, os.cpu_count() or 1), log=log)
b = [workload.load_fastq_fast(f) for f in fqs if f]
ix = Index(ixf, 0)
log("index resident: %.2f GB" % (ix.device_bytes / 1e9))
first, out = None, {}
import helpers
for v in variants:
    env = dict(kv.split("=", 1) for kv in v.split(",") if kv)
    for k, val in env.items():
        os.environ[k] = val
    m = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max([x.max_len for x in b] + [64]))
    m.upload(*b); m.sync()
    best = None
    for it in range(5):
        m.run(); m.sync()
        if it >= 2 and (best is None or m.last_kernel_ms < best[0]):
            best = (m.last_kernel_ms, list(m.last_seed_times) if m.binned else None, list(m.last_kernel_times))
    rec = {"ms": best[0], "reads_per_s": (2 if paired else 1) * b[0].n / best[0] * 1e3,
           "seeding_ms": dict(zip(m.SEED_KERNELS, best[1])) if best[1] else None, "phases_ms": dict(zip(m.KERNELS, best[2]))}
    if os.environ.get("SWEEP_E2E"):
        # end to end through abg_map_batch: page-locked host buffers in and out, copies inside the timed region
        if "bp" not in globals():
            from abismal_b200.capi import Results
            bp = [x.to_pinned() for x in b]
            rp = Results(b[0].n, paired, m.stride, pinned=True)
        m.map_batch(*bp, results=rp)
        t_e2e = []
        for _ in range(3):
            t0 = time.perf_counter()
            m.map_batch(*bp, results=rp)
            t_e2e.append(1e3 * (time.perf_counter() - t0))
        rec["e2e_ms"] = min(t_e2e)
        rec["e2e_reads_per_s"] = (2 if paired else 1) * b[0].n / min(t_e2e) * 1e3
    res = m.map_batch(*b)
    if first is None:
        first = res
        rec["same_as_first"] = True
    else:
        try:
            helpers.assert_results_equal(res, first, paired)
            rec["same_as_first"] = True
        except AssertionError as e:
            rec["same_as_first"] = False
            rec["diff"] = str(e)[:300]
    log("variant [%s] %s" % (v, json.dumps(rec)))
    out[v] = rec
    if check_n and v == variants[-1]:
        o = helpers.OracleMapper(ixf, mode=mode)
        want = o.map_batch(*[x.slice(0, check_n) for x in b])
        m2 = Mapper(ix, mode=mode, max_batch=check_n, max_read_len=max([x.max_len for x in b] + [64]))
        got = m2.map_batch(*[x.slice(0, check_n) for x in b])
        helpers.assert_results_equal(got, want, paired)
        log("parity vs oracle ok on %d items" % check_n)
        m2.close(); o.close()
    m.close()
    for k in env:
        del os.environ[k]
print(json.dumps(out))
