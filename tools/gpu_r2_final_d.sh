#!/bin/bash
# round 2, final batch d (one GPU), after the staged panel kernel: suite, smoke, default bench line, reference arm,
# then ncu --set full of the panel kernel in its default form on a 300 000 x 1000 cohort
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3
echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; echo ref rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:modcov_panel_kernel -c 1 -o gpurun_out/r2_prof_modcov_final python bench.py --workload cfg5 --targets 300000 --steps 1 --warmup 0 --quick > gpurun_out/r2_ncu_modcov_final.log 2>&1
echo ncu rc=$?
