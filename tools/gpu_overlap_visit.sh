#!/bin/bash
# GPU-box visit for the launch-shape study (tools/overlap_perf.py).  PYTEST=1 runs the parity tests first.
OUT=gpurun_out/${1:-r01_overlap}
CONFIGS=${2:-0,0:2,1,0:::4,0:::2}
mkdir -p $OUT
if [ -n "$PYTEST" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 900 python tools/overlap_perf.py 1048576 $CONFIGS 2000 > $OUT/op.log 2>&1
grep "config\|MISMATCH\|parity\|Error\|error" $OUT/op.log
