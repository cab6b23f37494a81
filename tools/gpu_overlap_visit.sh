#!/bin/bash
# GPU-box visit for the launch-shape study: sequential vs overlapped launch on four builds of the library
# (default = 4 cp.async.ca stages, cp.async.cg, 2 stages, no staging).  PYTEST=1 runs the parity tests first.
OUT=gpurun_out/${1:-r01_overlap}
mkdir -p $OUT
if [ -n "$PYTEST" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
run() {  # name, configs, [check_n]
  lib=$PWD/abismal_b200/libabismal_b200$1.so
  ABISMAL_B200_LIB=$lib timeout 600 python tools/overlap_perf.py 1048576 $2 $3 > $OUT/op$1.log 2>&1
  grep "config\|MISMATCH\|parity\|Error\|error" $OUT/op$1.log
}
run "_nostage" 0,1,1:2:1:3 2000
run "" 0,1 2000
run "_cg" 0
run "_s2" 0,1
