#!/bin/bash
CNVB_WIS_TRACE=1 timeout 600 python -m pytest tests/test_gpu_wis.py -x -q -k "matches_reference_sets and 5000-100-100-22-True" 2>&1 | tail -60
