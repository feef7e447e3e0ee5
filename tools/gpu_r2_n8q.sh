#!/bin/bash
# round 2: headline workload only, device-resident legs, on N GPUs (strong scaling of one genome): bash tools/gpu_r2_n8q.sh N
N=$1
for n in $N; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $n --steps 50 --warmup 5 --workload coverage --quick > gpurun_out/r2_cov_q_n$n.json 2> gpurun_out/r2_cov_q_n$n.err
echo rc=$?; grep "^{" gpurun_out/r2_cov_q_n$n.json | cut -c1-700
done
