#!/bin/bash
# Round-2 visit 10: tile-sorted scatter and static filter distribution against the defaults (full PBAT batch),
# then ncu --set full of the four seeding kernels on a quarter batch.
TAG=${1:-r02_v10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_GRAB=0;ABISMAL_B200_SCATTER_SORT=1;ABISMAL_B200_SCATTER_SORT=1,ABISMAL_B200_BIN_SHIFT=20;ABISMAL_B200_SCATTER_SORT=1,ABISMAL_B200_BIN_SHIFT=18;ABISMAL_B200_SCATTER_SORT=1,ABISMAL_B200_FILTER_GRAB=0;ABISMAL_B200_SCATTER_SORT=1,ABISMAL_B200_BIN_SHIFT=18,ABISMAL_B200_FILTER_GRAB=0" 4000 > $OUT/sweep_pbat.log 2>&1
echo "sweep pbat exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_pbat.log | cut -c1-560
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hash_kernel|scatter|filter_kernel|seed_kernel' -c 4 \
    -f -o $OUT/seeding_full python bench.py --pairs 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-cli > $OUT/full_bench.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/full_bench.log | cut -c1-200
ls -la $OUT
