#!/bin/bash
# Round-2 visit 9: binned seeding, fourth build (scatter writes tuples only, the filter builds the payloads from
# the strands' planes in shared memory): sweep of scatter CTAs / bin size / grab size / pipelined filter on the
# full PBAT batch, binned vs direct on the SE and random-PBAT batches.
TAG=${1:-r02_v9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "binned or (records_equal_oracle and not direct and not rows and not warp)" > $OUT/pytest_core.log 2>&1
echo "pytest core exit $?"; tail -4 $OUT/pytest_core.log
timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_PIPE=1;ABISMAL_B200_FILTER_GRAB=512;ABISMAL_B200_FILTER_GRAB=1024;ABISMAL_B200_SCATTER_CTAS=2;ABISMAL_B200_BIN_SHIFT=20;ABISMAL_B200_BIN_SHIFT=18;ABISMAL_B200_BIN_SHIFT=18,ABISMAL_B200_SCATTER_CTAS=2;ABISMAL_B200_BINS=0;ABISMAL_B200_FILTER_GRAB=256" 4000 > $OUT/sweep_pbat.log 2>&1
echo "sweep pbat exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_pbat.log | cut -c1-560
timeout 600 python tools/env_sweep.py se 1048576 "ABISMAL_B200_BINS=0;ABISMAL_B200_BINS=1" 4000 > $OUT/sweep_se.log 2>&1
echo "sweep se exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_se.log | cut -c1-560
timeout 600 python tools/env_sweep.py rpbat 1048576 "ABISMAL_B200_BINS=0;ABISMAL_B200_BINS=1" 4000 > $OUT/sweep_rpbat.log 2>&1
echo "sweep rpbat exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_rpbat.log | cut -c1-560
ls -la $OUT
