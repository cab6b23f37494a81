#!/bin/bash
# Round-2 visit 7: binned seeding, second build (filter v2, 32-byte payload stores, replay without a second encode):
# parity, bench, ncu --set full of the scatter and filter kernels at the full batch.
TAG=${1:-r02_v7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_parity.log 2>&1
echo "pytest parity exit $?"; tail -8 $OUT/pytest_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cli > $OUT/bench_pbat.json 2> $OUT/bench_pbat.log
echo "bench pbat exit $?"; tail -3 $OUT/bench_pbat.log
ABISMAL_B200_BINS=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cli > $OUT/bench_pbat_direct.json 2> $OUT/bench_pbat_direct.log
echo "bench pbat direct exit $?"
python - $TAG <<'PY'
import json,sys
for f in ("bench_pbat.json","bench_pbat_direct.json"):
    try:
        d=json.load(open("gpurun_out/%s/%s" % (sys.argv[1], f)))
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:round(v["ms_per_launch"],2) for k,v in d["kernels"].items()}, d.get("binned_seeding"), d["parity"]["mismatching_records"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scatter_kernel|filter_kernel' -c 2 \
    -f -o $OUT/scatter_filter_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cli > $OUT/full_bench.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/full_bench.log | cut -c1-200
ls -la $OUT
