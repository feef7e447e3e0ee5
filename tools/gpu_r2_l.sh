#!/bin/bash
# round 2, batch l: the whole GPU suite after the WIS rework, then config 5 (streamed sweep shares the epilogue changes)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/r2_cfg5_l.json 2> gpurun_out/r2_cfg5_l.err; echo cfg5 rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_cfg5_l.json'))
print('cfg5 step %.1f ms' % d['ms_per_step'], {n: round(v['ms_per_step'],1) for n,v in d['kernels'].items()}, d['roofline']['frac'], d['wis_stats'], d['checks'], d['clocks'])
PY
