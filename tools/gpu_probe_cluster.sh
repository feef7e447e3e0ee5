# guarded probe of the streamed sweep variants (every step under its own timeout)
timeout 200 python -m pytest tests/test_gpu_wis.py -m gpu -x -q 2>&1 | tail -2
for x in 0 1; do
  CNVB_WIS_XPOSE=$x timeout 150 python bench.py --workload cfg5 --targets 300000 --quick --steps 3 --warmup 2 > gpurun_out/xp_$x.json 2> gpurun_out/xp_$x.err
  python -c "
import json
d=json.load(open('gpurun_out/xp_$x.json')); print('xpose', $x, d['ms_per_step'], d['kernels']['wis_sweep_kernel'], d['roofline']['executed_tflops'], d['checks'])"
done
timeout 400 python bench.py --workload cfg5 --quick --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('full', d['ms_per_step'], d['kernels']['wis_sweep_kernel'], d['roofline']['executed_tflops'], d['checks'], d['clocks'])"
