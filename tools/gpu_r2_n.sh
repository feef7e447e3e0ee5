#!/bin/bash
# round 2: the default bench on N GPUs exactly as the driver launches it: bash tools/gpu_r2_n.sh N
N=$1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2>&1 | tail -3
echo rc=$?
tail -5 gpurun_out/r2_bench_n$N.err
