#!/bin/bash
# round 2, batch e (N GPUs): deferred region loop -- test, scale probe, quick strong-scaling lines
timeout 300 python -m pytest tests/test_gpu_run.py -x -q -k "deferred or region_loop" 2>&1 | tail -3
bash tools/gpu_scale_probe.sh "$1"
bash tools/gpu_r2_n8q.sh "$1"
