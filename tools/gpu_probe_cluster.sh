# guarded probe of the streamed sweep variants (every step under its own timeout)
timeout 150 python -m pytest tests/test_gpu_wis.py tests/test_gpu_run.py -m gpu -x -q 2>&1 | tail -2
for x in 0 1; do
  CNVB_WIS_XPOSE=$x timeout 150 python bench.py --workload cfg5 --targets 300000 --quick --steps 3 --warmup 2 > gpurun_out/xp_$x.json 2> gpurun_out/xp_$x.err
  echo "xpose=$x rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/xp_$x.json')); print($x, d['ms_per_step'], d['kernels']['wis_sweep_kernel'], d['roofline']['executed_tflops'], d['checks'], d['clocks']['sm_mhz'], d['wis_stats'])"
done
