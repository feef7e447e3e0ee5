#!/bin/bash
# round 2, batch f: new tests (files, indexed fetch, deferred loop) + ncu --set full of the cfg1 kernels
timeout 600 python -m pytest tests/test_gpu_files.py tests/test_bam.py tests/test_gpu_run.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest_f.log
timeout 300 python bench.py --workload cfg1 --steps 3 --warmup 1 --quick > gpurun_out/r2_cfg1_plain.json 2> gpurun_out/r2_cfg1_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'depth_tile_kernel|frag_bin_tgt_kernel|depth_prepare_kernel|window_finalize_kernel' -s 8 -c 8 -o gpurun_out/r2_prof_cfg1 python bench.py --workload cfg1 --steps 3 --warmup 1 --quick > gpurun_out/r2_ncu_cfg1.log 2>&1
echo ncu rc=$?
