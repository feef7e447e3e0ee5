#!/bin/bash
for n in $1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 tools/scale_probe.py 2> gpurun_out/r2_scale_probe_n$n.err | tee gpurun_out/r2_scale_probe_n$n.json
done
