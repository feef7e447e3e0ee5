#!/bin/bash
# ncu --set full of the other kernels of the config-3 step: the four compactions, the re-score, the calls
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'wis_compact_kernel|wis_rescore_kernel|wis_calls_kernel' -c 6 -o gpurun_out/r2_wis_rest python bench.py --workload wis --steps 1 --warmup 1 --quick > gpurun_out/r2_ncu_wis_rest.log 2>&1
echo ncu rc=$?
