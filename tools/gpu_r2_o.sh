#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_coverage.py tests/test_gpu_run.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('default step %.3f ms serialised %.3f' % (d['ms_per_step'], d.get('ms_per_step_serialised',0)), {n: (round(v['ms_per_step'],4), v['launches_per_step']) for n,v in d['kernels'].items()}, 'launches', d['gpu_launches'])"
