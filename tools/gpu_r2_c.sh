#!/bin/bash
# round 2, batch c: whole GPU suite + the default bench line + the reference arm
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_c.log
( time timeout 600 python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err ) 2>&1 | tail -3
echo bench rc=$?
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_c.json 2> gpurun_out/r2_bench_ref_c.err ) 2>&1 | tail -3
