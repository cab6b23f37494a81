#!/bin/bash
# Round-2 visit 2: the task-parallel alignment (enum_kernel -> dp_kernel): parity tests, A/B bench, launch list,
# front-end throughput on a large input.
TAG=${1:-r02_v2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cli > $OUT/bench_pbat.json 2> $OUT/bench_pbat.log
echo "bench pbat exit $?"; cat $OUT/bench_pbat.json; tail -5 $OUT/bench_pbat.log
ABISMAL_B200_TASKS=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_pbat_notasks.json 2> $OUT/bench_pbat_notasks.log
echo "bench pbat (no tasks) exit $?"; cat $OUT/bench_pbat_notasks.json
timeout 900 python bench.py --mode se --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_se.json 2> $OUT/bench_se.log
echo "bench se exit $?"; cat $OUT/bench_se.json
timeout 900 python bench.py --mode rpbat --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_rpbat.json 2> $OUT/bench_rpbat.log
echo "bench rpbat exit $?"; cat $OUT/bench_rpbat.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'seed_kernel|align_kernel|map_reads_kernel|enum_kernel|dp_kernel' \
    -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
timeout 900 python tools/cli_perf.py 1048576 8 200000 > $OUT/cli_perf.log 2>&1
echo "cli_perf exit $?"; tail -12 $OUT/cli_perf.log
ls -la $OUT
