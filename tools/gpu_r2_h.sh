#!/bin/bash
# round 2, batch h: per-round trace of the resident WIS sweep (config 3) -- tiles, candidates left by the sweep,
# launch times.  CNVB_WIS_DBG switches parts off for comparison (results are wrong there):
# 1 no TMEM reads / tests, 2 no MMAs, 4 no append stores, 8 no appends.
timeout 150 python -m pytest tests/test_gpu_wis.py -x -q 2>&1 | tail -3
for dbg in 0 1 8; do
echo "== CNVB_WIS_DBG=$dbg"
env CNVB_WIS_TRACE=1 CNVB_WIS_DBG=$dbg timeout 40 python bench.py --workload wis --steps 1 --warmup 1 --quick 2>&1 >/dev/null | grep "wis trace" | grep -v blocks | tail -8 | awk '{if ($9>0) printf "%s  -> %.0f clk/tile\n", $0, $6*1e-3*1.965e9*148/$9; else print $0}'
done
timeout 40 python bench.py --workload wis --steps 5 --warmup 2 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']; print('step %.2f ms' % d['ms_per_step'], {n: round(v['ms_per_step'],3) for n,v in k.items()}, d['roofline']['frac'], d['wis_stats'], d['checks'])"
