#!/bin/bash
# Round-2 visit 15: sub-batch sizes of the end-to-end path (abg_map_batch with host buffers).
TAG=${1:-r02_v15}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SWEEP_E2E=1 timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_CHUNK2=65536;ABISMAL_B200_CHUNK2=131072;ABISMAL_B200_CHUNK2=262144;ABISMAL_B200_CHUNK2=1048576;ABISMAL_B200_CHUNK=65536,ABISMAL_B200_CHUNK2=131072;ABISMAL_B200_CHUNK=131072,ABISMAL_B200_CHUNK2=262144;ABISMAL_B200_CHUNK=16384,ABISMAL_B200_CHUNK2=131072" 4000 > $OUT/sweep_e2e.log 2>&1
echo "sweep exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_e2e.log | cut -c1-900
ls -la $OUT
