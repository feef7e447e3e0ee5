#!/bin/bash
# round 2, batch g: split work items of the resident WIS sweep -- parity, then the cfg3 line with and without
timeout 900 python -m pytest tests/test_gpu_wis.py tests/test_gpu_files.py -x -q 2>&1 | tail -4
for sp in 1 0; do
CNVB_WIS_SPLIT=$sp timeout 300 python bench.py --workload wis --steps 5 --warmup 2 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']; print('split=$sp step %.2f ms' % d['ms_per_step'], {n: round(v['ms_per_step'],3) for n,v in k.items()}, d['roofline']['frac'], d['wis_stats'], d['checks'])"
done
