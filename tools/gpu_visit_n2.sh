#!/bin/bash
# Two-GPU visit: (a) bench.py under torch.distributed.run on 2 GPUs, (b) the front end with -gpus 2 against
# -gpus 1 on the same FASTQ files: SAM and statistics must be byte-identical (batches are sharded over the
# GPUs and written in input order; the statistics are summed across the GPUs' workers).
TAG=${1:-r02_n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/box.txt; nproc >> $OUT/box.txt; free -g >> $OUT/box.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.log
echo "bench n2 exit $?"; cut -c1-400 $OUT/bench_n2.json; tail -3 $OUT/bench_n2.log
D=/tmp/abismal_b200_bench/g3100000000_s20251017
CLI=abismal_b200/bin/abismal-b200
for g in 1 2; do
  $CLI map -v -P -gpus $g -t $(nproc) -i $D/genome.idx -o /dev/shm/cli_g$g.sam -s /dev/shm/cli_g$g.stats \
      $D/pbat_n1048576_r0_1.fq $D/pbat_n1048576_r0_2.fq > $OUT/cli_g$g.log 2>&1
  echo "cli -gpus $g exit $?"; grep "total mapping time\|stage busy\|index upload" $OUT/cli_g$g.log
done
grep -v "^@PG" /dev/shm/cli_g1.sam | md5sum > $OUT/sam_md5.txt; grep -v "^@PG" /dev/shm/cli_g2.sam | md5sum >> $OUT/sam_md5.txt
md5sum /dev/shm/cli_g1.stats /dev/shm/cli_g2.stats >> $OUT/sam_md5.txt
cat $OUT/sam_md5.txt
cp -r gpurun_out/bench_logs $OUT/ 2>/dev/null
ls -la $OUT
