#!/bin/bash
# Round-2 visit 6: row-parallel best_pair (parity with every pair forced through it, repeat-genome timing) and an
# ncu --set full capture of the binned seeding kernels.
TAG=${1:-r02_v6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_rows.log 2>&1
echo "pytest by_rows exit $?"; tail -15 $OUT/pytest_rows.log
timeout 600 python tools/repeat_perf.py 1e8 200000 20000 > $OUT/repeat_perf.log 2>&1
echo "repeat_perf exit $?"; grep "^\[rep\] tasks" $OUT/repeat_perf.log | cut -c1-900; tail -1 $OUT/repeat_perf.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hash_kernel|count_kernel|scatter_kernel|filter_kernel|seed_kernel' -c 5 \
    -f -o $OUT/seeding_full python bench.py --pairs 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-cli > $OUT/full_bench.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/full_bench.log | cut -c1-300
ls -la $OUT
