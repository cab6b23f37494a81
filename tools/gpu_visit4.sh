#!/bin/bash
# Round-2 visit 4: (a) microbenchmark of a binned prefilter (tools/micro/gather_bench4.cu), (b) the repeat genome
# with larger set-arena / task capacities and run statistics, (c) the reads of rank 6 of an 8-GPU run on one GPU.
TAG=${1:-r02_v4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 tools/micro/gather_bench4 > $OUT/gather_bench4.txt 2>&1
echo "gather_bench4 exit $?"; cat $OUT/gather_bench4.txt
timeout 600 python tools/repeat_perf.py 1e8 200000 20000 > $OUT/repeat_perf.log 2>&1
echo "repeat_perf exit $?"; grep "^\[rep\] tasks" $OUT/repeat_perf.log | cut -c1-900
timeout 600 python bench.py --as-rank 6 --steps 3 --warmup 3 --no-cpu-baseline --no-cli > $OUT/bench_rank6.json 2> $OUT/bench_rank6.log
echo "bench as-rank 6 exit $?"; cut -c1-300 $OUT/bench_rank6.json; tail -3 $OUT/bench_rank6.log
