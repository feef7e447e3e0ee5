#!/bin/bash
# round 2, final batch a (one GPU): suite, default bench line (all workloads), reference arm, launch lists,
# ncu --set full of the dominant kernels.  Each ncu pass only after the same command has exited 0 without it.
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_final_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3
echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; echo ref rc=$?
# launch lists
timeout 300 python bench.py --steps 2 --warmup 1 --quick > gpurun_out/r2_final_cov_q.json 2>/dev/null &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_cov_launches.csv python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1; echo launches cov rc=$?
timeout 300 python bench.py --workload wis --steps 1 --warmup 1 --quick > gpurun_out/r2_final_wis_q.json 2>/dev/null &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_final_wis_launches.csv python bench.py --workload wis --steps 1 --warmup 1 --quick > /dev/null 2>&1; echo launches wis rc=$?
# full captures: binning + refstats (default workload), the four sweep launches of config 3, the HMM kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'frag_bin_gw_kernel|refstats_lut_kernel|gc_radix_hist_kernel' -s 30 -c 6 -o gpurun_out/r2_final_prof_cov python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1; echo ncu cov rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wis_sweep_res_kernel' -c 4 -o gpurun_out/r2_final_prof_wis python bench.py --workload wis --steps 1 --warmup 1 --quick > /dev/null 2>&1; echo ncu wis rc=$?
timeout 300 python bench.py --workload hmm --steps 1 --warmup 1 --quick > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hmm_' -c 5 -o gpurun_out/r2_final_prof_hmm python bench.py --workload hmm --steps 1 --warmup 1 --quick > /dev/null 2>&1; echo ncu hmm rc=$?
