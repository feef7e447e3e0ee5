#!/bin/bash
# ncu --set full of the round-3 sweep launch (4th wis_sweep_kernel launch) for both tile forms
for x in 1 0; do
CNVB_WIS_XPOSE_SMALL=$x timeout 600 ncu --set full --clock-control none --import-source on -k regex:wis_sweep_kernel -s 3 -c 1 -o gpurun_out/r2_wis_r3_x$x python bench.py --workload wis --steps 1 --warmup 1 --quick > gpurun_out/r2_ncu_wis_x$x.log 2>&1
echo rc=$?
done
