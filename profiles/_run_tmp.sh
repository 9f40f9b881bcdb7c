timeout 600 python -m pytest tests/test_gpu_rootq_obs.py -q -x -k "host" 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --no-qat > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; tail -3 gpurun_out/r02_bench_n1_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_b.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d['e2e'], indent=1)[:1500])
print(d['cpu_baseline']['value'])
PY
