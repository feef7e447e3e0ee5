python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r1_bench_coverage.json 2> gpurun_out/r1_bench_coverage.err; echo rc=$?
python bench.py --workload wis --steps 10 --warmup 3 > gpurun_out/r1_bench_wis.json 2> gpurun_out/r1_bench_wis.err; echo rc=$?
python bench.py --workload cfg1 > gpurun_out/r1_bench_cfg1.json 2> gpurun_out/r1_bench_cfg1.err; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2>/dev/null; echo rc=$?
python -c "
import json
for f in ('coverage','wis','cfg1'):
    d=json.load(open('gpurun_out/r1_bench_%s.json'%f)); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('host_bam_decode'))
"
