#!/bin/bash
# Round-2 visit 1: parity tests, bench lines of the three modes (with the FASTQ->SAM leg and the SAM comparison
# against the reference binary), reproduction of the reads of ranks 4-7 of an 8-GPU run on one GPU, and a
# compute-sanitizer pass over a small batch.  Everything lands in gpurun_out/<tag>/.
TAG=${1:-r02_v1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
(nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; free -g; df -h /tmp /dev/shm; ulimit -a) > $OUT/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_pbat.json 2> $OUT/bench_pbat.log
echo "bench pbat exit $?"; cat $OUT/bench_pbat.json
for r in 6 4 5 7; do
  RANK=$r WORLD_SIZE=1 LOCAL_RANK=0 ABISMAL_B200_SIM_PROCS=4 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline \
     > $OUT/repro_rank$r.json 2> $OUT/repro_rank$r.log
  echo "repro rank $r exit $?"; tail -3 $OUT/repro_rank$r.log
done
timeout 900 python bench.py --mode se --steps 5 --warmup 3 > $OUT/bench_se.json 2> $OUT/bench_se.log
echo "bench se exit $?"; cat $OUT/bench_se.json
timeout 900 python bench.py --mode rpbat --steps 5 --warmup 3 > $OUT/bench_rpbat.json 2> $OUT/bench_rpbat.log
echo "bench rpbat exit $?"; cat $OUT/bench_rpbat.json
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 200 > $OUT/memcheck.log 2>&1
echo "memcheck exit $?"; tail -5 $OUT/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py 100 > $OUT/racecheck.log 2>&1
echo "racecheck exit $?"; tail -5 $OUT/racecheck.log
cp -r gpurun_out/bench_logs $OUT/ 2>/dev/null
ls -la $OUT
