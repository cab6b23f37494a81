#!/bin/bash
# Round-2 visit 11: new defaults (tile-sorted scatter, 32 MB bins, 40-byte plane rows): the whole GPU test suite,
# then L2 cache-policy variants of the filter and larger bins on the full PBAT batch.
TAG=${1:-r02_v11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -5 $OUT/pytest_gpu.log
timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_CACHE=1;ABISMAL_B200_FILTER_CACHE=2;ABISMAL_B200_FILTER_CACHE=3;ABISMAL_B200_BIN_SHIFT=21;ABISMAL_B200_BIN_SHIFT=22;ABISMAL_B200_BIN_SHIFT=21,ABISMAL_B200_FILTER_CACHE=3;ABISMAL_B200_FILTER_GRAB=128" 4000 > $OUT/sweep_pbat.log 2>&1
echo "sweep pbat exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_pbat.log | cut -c1-560
ls -la $OUT
