#!/bin/bash
# HMM matrix kernel with the rank-one shortcut: parity, then the config-4 line
timeout 600 python -m pytest tests/test_gpu_hmm.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload hmm --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('hmm step %.3f ms value %.3e frac %.3f' % (d['ms_per_step'], d['value'], d['roofline']['frac']), {n: round(v['ms_per_step'],4) for n,v in d['kernels'].items()}, d['checks'])"
