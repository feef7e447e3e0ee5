#!/bin/bash
# round 2, final batch e (one GPU), after the host decoder's own DEFLATE / CRC-32 and the threaded record walk:
# suite, smoke, default bench line (host_bam_decode and file_level legs re-taken on the box's 16 host threads)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3
echo bench rc=$?
