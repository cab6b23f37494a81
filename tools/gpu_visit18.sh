#!/bin/bash
# Round-2 visit 18: filter bounded for 8 CTAs per SM; timeline of abg_map_batch (where the end-to-end call spends
# the 13 ms it takes longer than the device-resident kernels).
TAG=${1:-r02_v18}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SWEEP_E2E=1 ABISMAL_B200_VERBOSE=1 timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_MINB=8;ABISMAL_B200_FILTER_MINB=8,ABISMAL_B200_FILTER_GRAB=128;ABISMAL_B200_CHUNK=65536" 4000 > $OUT/sweep.log 2>&1
echo "sweep exit $?"; grep "variant\|parity\|Error\|error\|abg_map_batch\]" $OUT/sweep.log | cut -c1-600
ls -la $OUT
