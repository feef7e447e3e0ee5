#!/bin/bash
# round 2, final batch b (one GPU), after the HMM change: suite, smoke, default bench line, reference arm, the launch
# list of the default workload restricted to the library's kernels (batch a's 400 launches were used up by the
# generation of the synthetic genome), ncu --set full of the HMM kernels.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3
echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; echo ref rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'frag_bin|refstats|gc_radix|norm_|cov_records|window_stats|target_stats|sum_' -c 400 --csv --log-file gpurun_out/r2_final_cov_launches.csv python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1; echo launches cov rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hmm_' -c 5 -o gpurun_out/r2_final_prof_hmm python bench.py --workload hmm --steps 1 --warmup 1 --quick > /dev/null 2>&1; echo ncu hmm rc=$?
