#!/bin/bash
# Round-2 visit 13: cp.async-staged filter, register bounds of the selection / enumeration / replay kernels,
# variant tests, where the front end's index load time goes.
TAG=${1:-r02_v13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants" > $OUT/pytest_variants.log 2>&1
echo "pytest variants exit $?"; tail -4 $OUT/pytest_variants.log
timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_PIPE=1;ABISMAL_B200_FILTER_PIPE=1,ABISMAL_B200_FILTER_GRAB=128;ABISMAL_B200_MINB_ALIGN=3;ABISMAL_B200_MINB_ALIGN=2;ABISMAL_B200_MINB_ENUM=3;ABISMAL_B200_MINB_SEED=3;ABISMAL_B200_MINB_SEED=5" 4000 > $OUT/sweep_pbat.log 2>&1
echo "sweep pbat exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_pbat.log | cut -c1-700
D=/tmp/abismal_b200_bench/g3100000000_s20251017
for k in 1 2; do
ABISMAL_B200_VERBOSE=1 abismal_b200/bin/abismal-b200 map -v -P -t 16 -i $D/genome.idx -o /dev/shm/x.sam $D/pbat_n1048576_r0_1.fq $D/pbat_n1048576_r0_2.fq > $OUT/cli_verbose_$k.log 2>&1
echo "cli exit $?"; grep "abg_index_create\|total mapping\|stage busy\|index upload\|loading" $OUT/cli_verbose_$k.log
done
ls -la $OUT
