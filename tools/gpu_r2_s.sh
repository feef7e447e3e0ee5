#!/bin/bash
# strong scaling of the default workload on N GPUs, device-resident legs: planned (CUDA graph) vs eager region loop
N=$1
for g in "" "--no-graph"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 50 --warmup 5 --workload coverage --quick $g > gpurun_out/r2_cov_q_n${N}${g}.json 2> gpurun_out/r2_cov_q_n${N}${g}.err
echo rc=$?; grep "^{" gpurun_out/r2_cov_q_n${N}${g}.json | cut -c1-330
done
