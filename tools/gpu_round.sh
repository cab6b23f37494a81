#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the bench command and one
# `ncu --set full` capture of the two-phase kernels.  Everything lands in gpurun_out/<tag>/.
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.log
echo "bench exit $?"; cat $OUT/bench.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.log
  cat $OUT/bench_reference.json
fi
if [ -x tools/micro/gather_bench3 ]; then
  # random-gather ceiling of this box: 8/32/64/128-byte units over a 3 GB and a 20 GB array
  (timeout 120 tools/micro/gather_bench3 3e9; timeout 120 tools/micro/gather_bench3 2e10) > $OUT/gather_bench3.txt 2>&1
fi
if [ -z "$SKIP_NCU" ]; then
  # launch list of the bench command (cold-cache, serialised: shares must agree, not absolutes)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'seed_kernel|align_kernel|map_reads_kernel' \
    -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
  # full capture of one seed_kernel + one align_kernel launch on a quarter-size batch of the same reads
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'seed_kernel|align_kernel' -c 2 \
    -f -o $OUT/two_phase_full python bench.py --pairs 262144 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/full_bench.log 2>&1
  ls -la $OUT
fi
