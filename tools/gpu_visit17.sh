#!/bin/bash
# Round-2 visit 17: resident filter CTAs (the live window of the record array against latency hiding), with and
# without the cp.async-staged records; the repeat-rich genome with the reference binary next to it.
TAG=${1:-r02_v17}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/env_sweep.py pbat 1048576 ";ABISMAL_B200_FILTER_CTAS=4;ABISMAL_B200_FILTER_CTAS=3;ABISMAL_B200_FILTER_CTAS=2;ABISMAL_B200_FILTER_CTAS=3,ABISMAL_B200_FILTER_PIPE=1;ABISMAL_B200_FILTER_CTAS=2,ABISMAL_B200_FILTER_PIPE=1;ABISMAL_B200_FILTER_CTAS=3,ABISMAL_B200_FILTER_PIPE=1,ABISMAL_B200_FILTER_GRAB=128;ABISMAL_B200_FILTER_CTAS=4,ABISMAL_B200_FILTER_PIPE=1,ABISMAL_B200_FILTER_GRAB=128;ABISMAL_B200_FILTER_CTAS=2,ABISMAL_B200_FILTER_PIPE=1,ABISMAL_B200_FILTER_GRAB=128;ABISMAL_B200_FILTER_CTAS=3,ABISMAL_B200_BIN_SHIFT=19" 4000 > $OUT/sweep_pbat.log 2>&1
echo "sweep exit $?"; grep "variant\|parity\|Error\|error" $OUT/sweep_pbat.log | cut -c1-420
timeout 600 python tools/repeat_perf.py 1e8 200000 100000 > $OUT/repeat_perf.log 2>&1
echo "repeat_perf exit $?"; tail -1 $OUT/repeat_perf.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('reference'))"
ls -la $OUT
