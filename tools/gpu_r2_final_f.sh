#!/bin/bash
# round 2, final batch f (one GPU): suite with the final host decoder, then its two bench legs on the box's host threads
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r2_final_pytest.log
timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/r2_host_decode_legs.json
import importlib.util, json, sys
sys.argv = ["bench.py"]
spec = importlib.util.spec_from_file_location("bench", "bench.py"); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
print(json.dumps({"host_bam_decode": m.host_bam_decode_leg(), "host_bam_decode_realistic": m.host_bam_decode_leg(n_pairs=200_000, payload="random")}))
PY
