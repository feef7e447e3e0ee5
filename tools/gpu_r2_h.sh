#!/bin/bash
for dbg in 1 17; do
echo "== RES=0 CNVB_WIS_DBG=$dbg"
env CNVB_WIS_RES=0 CNVB_WIS_TRACE=1 CNVB_WIS_DBG=$dbg timeout 40 python bench.py --workload wis --steps 1 --warmup 1 --quick 2>&1 >/dev/null | grep "wis trace" | grep sweep | tail -5 | awk '{if ($9>0) printf "%s  -> %.0f clk/tile\n", $0, $6*1e-3*1.965e9*148/$9; else print $0}'
done
