#!/bin/bash
# round 2: staged modcov_panel_kernel -- suite, then config 5 (device-resident legs) in the forms of the kernel
# (CNVB_PANEL_STAGE: 3 = 16-byte cp.async, the default; 2 = bulk copies; 1 = 4-byte cp.async; 0 = unstaged)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_w_pytest.log
for v in ${PANEL_FORMS:-"3 2" "2 2"}; do
set -- $v
CNVB_PANEL_STAGE=$1 CNVB_PANEL_SPL=$2 timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 1 --quick > gpurun_out/r2_w_cfg5_stage$1_spl$2.json 2> gpurun_out/r2_w_cfg5_stage$1_spl$2.err
echo stage=$1 spl=$2 rc=$?; grep "^{" gpurun_out/r2_w_cfg5_stage$1_spl$2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels'].get('modcov_panel_kernel'), d['clocks']['sm_mhz'], d['checks'])"
done
