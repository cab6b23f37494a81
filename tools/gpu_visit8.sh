#!/bin/bash
# Round-2 visit 8: binned seeding, third build (scatter: 2 CTAs/SM + next tuple in flight + interleaved planes;
# replay: only survivors a phase could still accept are ranked and replayed; hash: tuple cursor in registers;
# filter: grab size at run time): core parity, a sweep of the grab size / bin size on the full batch, CLI front end.
TAG=${1:-r02_v8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "records_equal_oracle or binned or batch_split or empty_and_ragged or pipelined" > $OUT/pytest_core.log 2>&1
echo "pytest core exit $?"; tail -4 $OUT/pytest_core.log
timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_GRAB=32;ABISMAL_B200_FILTER_GRAB=128;ABISMAL_B200_FILTER_GRAB=256;ABISMAL_B200_BIN_SHIFT=18;ABISMAL_B200_BIN_SHIFT=18,ABISMAL_B200_FILTER_GRAB=32;ABISMAL_B200_BIN_SHIFT=20;ABISMAL_B200_BINS=0;ABISMAL_B200_FILTER_GRAB=64" 4000 > $OUT/sweep.log 2>&1
echo "sweep exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep.log | cut -c1-700
timeout 600 python tools/cli_perf.py > $OUT/cli_perf.log 2>&1
echo "cli_perf exit $?"; grep "^\[cli\]" $OUT/cli_perf.log | cut -c1-420
ls -la $OUT
