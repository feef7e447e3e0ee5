#!/bin/bash
# round 2: config 5 on 2 GPUs with the staged panel kernel, and the multi-GPU parity tool of the row-sharded model
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --workload cfg5 --gpus $N --steps 2 --warmup 1 --quick > gpurun_out/r2_bench_cfg5_n$N.json 2> gpurun_out/r2_bench_cfg5_n$N.err
echo rc=$?
grep "rror" gpurun_out/r2_bench_cfg5_n$N.err | tail -3
python - <<PY
import json
for l in open("gpurun_out/r2_bench_cfg5_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["ms_per_step"], d["value"], d["stages_s"], d["checks"], d["clocks"]["sm_mhz"], d["kernels"].get("modcov_panel_kernel"))
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/check_wis_distributed.py 2>&1 | tail -3
