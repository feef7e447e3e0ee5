#!/bin/bash
# ncu --set full of modcov_panel_kernel inside a config-5 chain of 300 000 targets x 1000 samples (one launch; the
# panel of 1.2 GB still exceeds L2 ten times over, the replay passes of the full-size launch took 4.5 minutes)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:modcov_panel_kernel -c 1 -o gpurun_out/r2_prof_modcov_staged python bench.py --workload cfg5 --targets 300000 --steps 1 --warmup 0 --quick > gpurun_out/r2_ncu_modcov_staged.log 2>&1
echo rc=$?
