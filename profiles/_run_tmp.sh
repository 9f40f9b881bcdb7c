python -m pytest tests/test_gpu_bnq.py tests/test_gpu_parity_edges.py tests/test_gpu_modules.py -q -x 2>&1 | tail -4
python bench.py --qat-only --qat-arms small_batch 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())['qat_images_per_s']
for f in d:
    if isinstance(d[f], dict) and f not in ('arms',): print(f, {k:((v.get('images_per_s'), v.get('ms_per_step'), v.get('error')) if isinstance(v, dict) else v) for k,v in d[f].items()})
"
