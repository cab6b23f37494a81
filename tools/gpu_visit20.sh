#!/bin/bash
# Round-2 visit 20: (a) the GPU index builder against `abismal idx` of the reference at 100 Mbp (30 % repeats,
# three-letter entries present): byte identity; (b) the PBAT bench line of the final build.
TAG=${1:-r02_v20}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/repeat_perf.py 1e8 20000 5000 > $OUT/repeat_small.log 2>&1; echo "repeat genome + GPU index exit $?"
D=/tmp/abismal_b200_bench/repeat_100000000
( time oracle/_ref/abismal idx -t $(nproc) $D/g.fa $D/ref.idx ) > $OUT/ref_idx.log 2>&1; echo "reference idx exit $?"; tail -4 $OUT/ref_idx.log
( time abismal_b200/bin/abismal-b200 idx $D/g.fa $D/ours.idx ) > $OUT/ours_idx.log 2>&1; echo "gpu idx exit $?"; tail -4 $OUT/ours_idx.log
ls -l $D/ref.idx $D/ours.idx $D/g.idx | tee $OUT/idx_files.txt
md5sum $D/ref.idx $D/ours.idx $D/g.idx | tee -a $OUT/idx_files.txt
cmp $D/ref.idx $D/ours.idx && echo "IDENTICAL: GPU builder == abismal idx at 100 Mbp" | tee -a $OUT/idx_files.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_pbat.json 2> $OUT/bench_pbat.log
echo "bench pbat exit $?"; python - $OUT/bench_pbat.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['fastq_to_sam']['value'], d['fastq_to_sam']['seconds'], d['parity']['mismatching_records'], d['parity']['reference_binary']['mismatching_records'], d['roofline']['frac'], d['roofline']['traffic'])
PY
ls -la $OUT
