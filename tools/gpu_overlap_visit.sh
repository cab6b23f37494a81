#!/bin/bash
# GPU-box visit for the launch-shape study: parity tests on the default build, then sequential vs overlapped
# launch on three builds of the library (default = 4 cp.async stages, no staging, 2 stages).
OUT=gpurun_out/${1:-r01_overlap}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python tools/overlap_perf.py 1048576 0,1,1:2:1:3,1:1:1,1:1:2 2000 > $OUT/op_default.log 2>&1
grep "config\|MISMATCH\|parity\|Error\|error" $OUT/op_default.log
ABISMAL_B200_LIB=$PWD/abismal_b200/libabismal_b200_nostage.so timeout 600 python tools/overlap_perf.py 1048576 0,1,1:1:2 > $OUT/op_nostage.log 2>&1
grep "config\|MISMATCH\|parity\|Error\|error" $OUT/op_nostage.log
ABISMAL_B200_LIB=$PWD/abismal_b200/libabismal_b200_s2.so timeout 600 python tools/overlap_perf.py 1048576 0,1,1:3:1:4 > $OUT/op_s2.log 2>&1
grep "config\|MISMATCH\|parity\|Error\|error" $OUT/op_s2.log
