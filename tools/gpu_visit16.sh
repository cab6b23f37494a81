#!/bin/bash
# Round-2 visit 16: the front end on the repeat-rich genome (tools/repeat_perf.py died in `abismal-b200 map`
# on its 100 000-pair sample): the error text, and which capacity / path it depends on.
TAG=${1:-r02_v16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tools/repeat_perf.py 1e8 200000 100000 > $OUT/repeat_perf.log 2>&1
echo "repeat_perf exit $?"; tail -3 $OUT/repeat_perf.log | cut -c1-300
D=/tmp/abismal_b200_bench/repeat_100000000
CLI=abismal_b200/bin/abismal-b200
run() {  # tag, env..., then the command's stderr tail
  tag=$1; shift
  env "$@" $CLI map -v -i $D/g.idx -o $D/ours_$tag.sam $D/pe_s_1.fq $D/pe_s_2.fq > $OUT/cli_$tag.log 2>&1
  echo "cli [$tag] exit $?"; tail -4 $OUT/cli_$tag.log | cut -c1-300
}
run default X=1
run scale4 ABISMAL_B200_TASK_SCALE=4
run ovf4096 ABISMAL_B200_OVF_PER_ITEM=4096
run notasks ABISMAL_B200_TASKS=0
run nobins ABISMAL_B200_BINS=0
run bigchunk ABISMAL_B200_CHUNK=262144
run workers1 X=1
$CLI map -v -gpu-workers 1 -i $D/g.idx -o $D/ours_w1.sam $D/pe_s_1.fq $D/pe_s_2.fq > $OUT/cli_w1.log 2>&1; echo "cli [w1] exit $?"; tail -3 $OUT/cli_w1.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 $CLI map -gpu-workers 1 -i $D/g.idx -o $D/ours_mc.sam $D/pe_s_1.fq $D/pe_s_2.fq > $OUT/memcheck.log 2>&1
echo "memcheck exit $?"; grep -m 12 "Invalid\|at \|ERROR SUMMARY\|Error" $OUT/memcheck.log | cut -c1-250
timeout 120 oracle/_ref/abismal map -t 16 -i $D/g.idx -o $D/ref.sam $D/pe_s_1.fq $D/pe_s_2.fq 2> /dev/null; echo "ref exit $?"
for t in default scale4 nobins w1; do
  if [ -s $D/ours_$t.sam ]; then python - $D/ref.sam $D/ours_$t.sam $t <<'PY'
import sys
a=sorted(l for l in open(sys.argv[1],'rb') if not l.startswith(b'@PG')); b=sorted(l for l in open(sys.argv[2],'rb') if not l.startswith(b'@PG'))
print(sys.argv[3], "records", len(a), len(b), "identical", a==b)
PY
  fi
done
ls -la $OUT
