#!/bin/bash
# Round-2 final record: whole GPU test suite, ncu launch list of the bench command, ncu --set full of every kernel
# of one full batch (-> profiles/traffic.json), bench lines of the three modes (the PBAT line with the front-end
# leg and the reference-binary SAM comparison), the reference arm, the repeat-rich genome.
TAG=${1:-r02_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
(nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; free -g) > $OUT/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cli > $OUT/launches_bench.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none \
    -k regex:'hash_kernel|count_kernel|bin_prefix_kernel|scatter_sorted_kernel|filter_kernel|seed_kernel|enum_kernel|dp_kernel|align_kernel' -c 9 \
    -f -o $OUT/all_kernels_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cli > $OUT/full_bench.log 2>&1
echo "ncu full exit $?"; tail -2 $OUT/full_bench.log | cut -c1-200
python tools/make_traffic.py $OUT/all_kernels_full.ncu-rep 1048576 pbat > $OUT/traffic_pbat.txt 2>&1; echo "traffic exit $?"; cp profiles/traffic.json $OUT/traffic.json
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_pbat.json 2> $OUT/bench_pbat.log
echo "bench pbat exit $?"; cut -c1-300 $OUT/bench_pbat.json; tail -2 $OUT/bench_pbat.log
timeout 900 python bench.py --mode se --steps 5 --warmup 3 --no-cli > $OUT/bench_se.json 2> $OUT/bench_se.log
echo "bench se exit $?"; cut -c1-300 $OUT/bench_se.json
timeout 900 python bench.py --mode rpbat --steps 5 --warmup 3 --no-cli > $OUT/bench_rpbat.json 2> $OUT/bench_rpbat.log
echo "bench rpbat exit $?"; cut -c1-300 $OUT/bench_rpbat.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.log
echo "bench reference exit $?"; cut -c1-300 $OUT/bench_reference.json
timeout 900 python tools/repeat_perf.py 1e8 200000 100000 > $OUT/repeat_perf.log 2>&1
echo "repeat_perf exit $?"; grep "^\[rep\] tasks" $OUT/repeat_perf.log | cut -c1-1200; tail -1 $OUT/repeat_perf.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('reference'))"
ls -la $OUT
