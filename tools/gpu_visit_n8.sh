#!/bin/bash
# Eight-GPU visit: bench.py under torch.distributed.run on 8 GPUs (every rank checks its own batch against the
# oracle; per-rank failures land in gpurun_out/bench_logs).
TAG=${1:-r02_n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
(nvidia-smi --query-gpu=index,name --format=csv; nproc; free -g; df -h /tmp) > $OUT/box.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/bench_n8.json 2> $OUT/bench_n8.log
echo "bench n8 exit $?"; cut -c1-600 $OUT/bench_n8.json; tail -5 $OUT/bench_n8.log | cut -c1-300
cp -r gpurun_out/bench_logs $OUT/ 2>/dev/null
ls -la $OUT $OUT/bench_logs 2>/dev/null
