#!/bin/bash
# Round-2 visit 5: first run of the binned seeding (hash -> scatter -> filter -> seed_kernel on the survivors).
TAG=${1:-r02_v5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/sanitize_small.py 400 > $OUT/small.log 2>&1
echo "small exit $?"; tail -8 $OUT/small.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_parity.log 2>&1
echo "pytest parity exit $?"; tail -30 $OUT/pytest_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_pbat.json 2> $OUT/bench_pbat.log
echo "bench pbat exit $?"; cat $OUT/bench_pbat.json | cut -c1-3000; tail -5 $OUT/bench_pbat.log
ABISMAL_B200_BIN_SHIFT=18 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_pbat_s18.json 2> $OUT/bench_pbat_s18.log
echo "bench pbat shift 18 exit $?"; python - <<'PY'
import json,sys
for f in ("bench_pbat.json","bench_pbat_s18.json"):
    try:
        d=json.load(open("gpurun_out/%s/%s" % (sys.argv[1] if len(sys.argv)>1 else "r02_v5", f)))
        print(f, d["value"], d["e2e"]["value"], {k:round(v["ms_per_launch"],2) for k,v in d["kernels"].items()}, d.get("binned_seeding"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 100 > $OUT/memcheck.log 2>&1
echo "memcheck exit $?"; tail -4 $OUT/memcheck.log
