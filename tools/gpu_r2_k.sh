#!/bin/bash
timeout 150 python -m pytest tests/test_gpu_wis.py -x -q 2>&1 | tail -3
for dbg in 0 1 8; do
echo "== CNVB_WIS_DBG=$dbg"
env CNVB_WIS_TRACE=1 CNVB_WIS_DBG=$dbg timeout 40 python bench.py --workload wis --steps 1 --warmup 1 --quick 2>&1 >/dev/null | grep "wis trace" | grep -v blocks | grep sweep | tail -4 | awk '{if ($9>0) printf "%s  -> %.0f clk/tile\n", $0, $6*1e-3*1.965e9*148/$9; else print $0}'
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wis_sweep_res_kernel -s 2 -c 1 -o gpurun_out/r2_wis_res_r2 python bench.py --workload wis --steps 1 --warmup 1 --quick > gpurun_out/r2_ncu_wis_res.log 2>&1
echo ncu rc=$?
