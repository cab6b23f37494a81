#!/bin/bash
# Round-2 visit 12: interleaved work cursors for the filter (small pieces without the one-address atomic limit),
# the variant tests, the front end with the multi-threaded index upload.
TAG=${1:-r02_v12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants" > $OUT/pytest_variants.log 2>&1
echo "pytest variants exit $?"; tail -4 $OUT/pytest_variants.log
timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_CURSORS=16,ABISMAL_B200_FILTER_GRAB=32;ABISMAL_B200_FILTER_CURSORS=16,ABISMAL_B200_FILTER_GRAB=64;ABISMAL_B200_FILTER_CURSORS=64,ABISMAL_B200_FILTER_GRAB=32;ABISMAL_B200_FILTER_CURSORS=8,ABISMAL_B200_FILTER_GRAB=64;ABISMAL_B200_FILTER_CURSORS=4,ABISMAL_B200_FILTER_GRAB=128;ABISMAL_B200_FILTER_CURSORS=16,ABISMAL_B200_FILTER_GRAB=32,ABISMAL_B200_BIN_SHIFT=19;ABISMAL_B200_FILTER_CURSORS=64,ABISMAL_B200_FILTER_GRAB=32,ABISMAL_B200_BIN_SHIFT=18" 4000 > $OUT/sweep_pbat.log 2>&1
echo "sweep pbat exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_pbat.log | cut -c1-560
timeout 600 python tools/cli_perf.py > $OUT/cli_perf.log 2>&1
echo "cli_perf exit $?"; grep "^\[cli\]" $OUT/cli_perf.log | cut -c1-420
ls -la $OUT
