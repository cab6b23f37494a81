#!/bin/bash
# Round-2 visit 19: sub-batch size of abg_map_batch with the host-side CIGAR scatter split over threads.
TAG=${1:-r02_v19}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SWEEP_E2E=1 timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_CHUNK=49152;ABISMAL_B200_CHUNK=65536;ABISMAL_B200_CHUNK=98304;ABISMAL_B200_CHUNK=131072;ABISMAL_B200_CHUNK=174763" 4000 > $OUT/sweep.log 2>&1
echo "sweep exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep.log | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pipelined or batch_split or empty_and_ragged" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest.log
ls -la $OUT
