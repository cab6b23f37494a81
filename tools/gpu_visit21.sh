#!/bin/bash
# Round-2 visit 21: (a) GPU index builder vs the reference's `abismal idx` at 100 Mbp (the reference's -t 16 run
# segfaults on this genome: fewer threads), (b) where the front end's time goes on 2^20 pairs (engine set-up).
TAG=${1:-r02_v21}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/repeat_perf.py 1e8 20000 5000 > $OUT/repeat_small.log 2>&1; echo "repeat genome + GPU index exit $?"
D=/tmp/abismal_b200_bench/repeat_100000000
for t in 8 4 2 1; do
  ( time oracle/_ref/abismal idx -t $t $D/g.fa $D/ref.idx ) > $OUT/ref_idx_t$t.log 2>&1; rc=$?
  echo "reference idx -t $t exit $rc"; tail -3 $OUT/ref_idx_t$t.log | tr '\n' ' '; echo
  if [ $rc -eq 0 ]; then break; fi
done
ls -l $D/ref.idx $D/g.idx | tee $OUT/idx_files.txt
md5sum $D/ref.idx $D/g.idx | tee -a $OUT/idx_files.txt
cmp $D/ref.idx $D/g.idx && echo "IDENTICAL: GPU builder == abismal idx at 100 Mbp" | tee -a $OUT/idx_files.txt
timeout 600 python tools/cli_perf.py 1048576 1 > $OUT/cli_perf.log 2>&1
echo "cli_perf exit $?"; grep "^\[cli\]" $OUT/cli_perf.log | cut -c1-460
ls -la $OUT
