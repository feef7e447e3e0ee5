#!/bin/bash
# round 2, batch b: WIS sweep phase timings, coverage kernels after the depth / targets changes
bash tools/wis_probe.sh 2>&1 | tee gpurun_out/r2_wis_probe.txt
timeout 600 python -m pytest tests/test_gpu_coverage.py tests/test_gpu_run.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --quick > gpurun_out/r2_cfg1_b.json 2> gpurun_out/r2_cfg1_b.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_cfg1_b.json"))
for k,v in d["kinds"].items():
    print(k, "call %.1f us" % (v["ms_per_step"]*1e3), "kernel", v["roofline"]["kernel"], "%.1f us" % (v["roofline"]["kernel_ms"]*1e3), "frac %.3f" % v["roofline"]["frac"], "hbm-only frac %.3f" % v["roofline"]["hbm_only"]["frac"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_cfg1_launches.csv python bench.py --workload cfg1 --steps 2 --warmup 1 --quick > /dev/null 2>&1
grep -c . gpurun_out/r2_cfg1_launches.csv
