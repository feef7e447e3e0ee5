# guarded probe of the cluster-multicast sweep (every step under its own timeout)
for c in 2 4; do
  CNVB_WIS_CLUSTER=$c timeout 150 python -m pytest tests/test_gpu_wis.py -m gpu -x -q 2>&1 | tail -2
done
for c in 2 4; do
  CNVB_WIS_CLUSTER=$c timeout 150 python bench.py --workload cfg5 --targets 300000 --quick --steps 3 --warmup 2 > gpurun_out/cl_$c.json 2> gpurun_out/cl_$c.err
  echo "cluster=$c rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/cl_$c.json')); print($c, d['ms_per_step'], d['kernels']['wis_sweep_kernel'], d['roofline']['executed_tflops'], d['checks'], d['clocks']['sm_mhz'])"
done
