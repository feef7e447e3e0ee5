# guarded probe of the sweep variants (every step under its own timeout)
timeout 120 python -m pytest tests/test_gpu_wis.py -m gpu -x -q 2>&1 | tail -2
CNVB_WIS_XPOSE_SMALL=1 timeout 120 python -m pytest tests/test_gpu_wis.py -m gpu -x -q 2>&1 | tail -2
for x in 0 1; do
  CNVB_WIS_XPOSE_SMALL=$x timeout 100 python bench.py --workload wis --quick --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('xpose_small', $x, d['ms_per_step'], d['kernels']['wis_sweep_kernel'], d['roofline']['executed_tflops'], d['wis_stats'])"
done
