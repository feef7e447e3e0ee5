#!/bin/bash
# timing experiments on the S <= 128 sweep (results under CNVB_WIS_DBG are wrong by construction)
for n256 in 1 0; do for dbg in 0 1 2 3; do
  CNVB_WIS_N256=$n256 CNVB_WIS_DBG=$dbg timeout 300 python bench.py --workload wis --steps 3 --warmup 2 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
sw=d['kernels']['wis_sweep_kernel']['ms_per_step']; t=d['wis_stats']['tiles']
print('n256=$n256 dbg=$dbg sweep %.3f ms tiles %d -> %.0f SM-clk per 256x128 tile; step %.2f ms' % (sw, t, sw*1e-3*1.965e9*148/max(t,1), d['ms_per_step']))"
done; done
