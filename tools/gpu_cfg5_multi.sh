# config 5 on N GPUs of one node (strong scaling): bash tools/gpu_cfg5_multi.sh N
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --workload cfg5 --gpus $N --steps 3 --warmup 2 --quick > gpurun_out/r1_bench_cfg5_n$N.json 2> gpurun_out/r1_bench_cfg5_n$N.err
echo rc=$?
grep "cfg5\|rror" gpurun_out/r1_bench_cfg5_n$N.err | tail -3
python - <<PY
import json
for l in open("gpurun_out/r1_bench_cfg5_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["ms_per_step"], d["value"], d["stages_s"], d["checks"], d["clocks"], d["roofline"]["executed_tflops"])
PY
