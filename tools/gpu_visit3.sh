#!/bin/bash
# Round-2 visit 3: overflow arena for large candidate sets + single-lane heap updates: parity tests, racecheck,
# the repeat-genome measurement, ncu full captures of the post-seeding kernels.
TAG=${1:-r02_v3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py 60 > $OUT/racecheck.log 2>&1
echo "racecheck exit $?"; grep -c "Race reported" $OUT/racecheck.log; tail -4 $OUT/racecheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 200 > $OUT/memcheck.log 2>&1
echo "memcheck exit $?"; tail -3 $OUT/memcheck.log
timeout 900 python tools/repeat_perf.py 1e8 200000 20000 > $OUT/repeat_perf.log 2>&1
echo "repeat_perf exit $?"; tail -6 $OUT/repeat_perf.log | cut -c1-1500
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_pbat.json 2> $OUT/bench_pbat.log
echo "bench pbat exit $?"; cat $OUT/bench_pbat.json | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'enum_kernel|dp_kernel|align_kernel' -c 3 \
    -f -o $OUT/post_seed_full python bench.py --pairs 262144 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/full_bench.log 2>&1
ls -la $OUT
