#!/bin/bash
# Eight-GPU visit, BASELINE configs[4]: random-PBAT pairs mapped with -R on 8 GPUs.
TAG=${1:-r02_n8_rpbat}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --mode rpbat --steps 5 --warmup 3 > $OUT/bench_n8_rpbat.json 2> $OUT/bench_n8_rpbat.log
echo "bench n8 rpbat exit $?"; tail -1 $OUT/bench_n8_rpbat.json | cut -c1-400; tail -4 $OUT/bench_n8_rpbat.log | cut -c1-300
mkdir -p gpurun_out/bench_logs; cp -r gpurun_out/bench_logs $OUT/
ls -la $OUT
