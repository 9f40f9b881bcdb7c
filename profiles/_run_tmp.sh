python -m pytest tests/test_gpu_bnq.py -q -x -k "not rewired" 2>&1 | tail -2
for v in "" "DLMCQ_BNQ_TILES=1 DLMCQ_BNQ_TILES_LIGHT=1" "DLMCQ_BNQ_TILES=2 DLMCQ_BNQ_TILES_LIGHT=2" "DLMCQ_BNQ_TILES=4 DLMCQ_BNQ_TILES_LIGHT=4" "DLMCQ_BNQ_LIGHT=0"; do echo "== $v"; env $v python profiles/prof_bnq.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['shape'], d['dtype'][:4], 'res' if d['residual'] else 'pla', 'fwd', d['fwd_us'], d['fwd_frac'], 'bwd', d['bwd_us'], d['bwd_frac'])
"; done
for v in "" "DLMCQ_BNQ_TILES=1 DLMCQ_BNQ_TILES_LIGHT=1"; do env $v python bench.py --qat-only --qat-arms ours_fused 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())['qat_images_per_s']
for f in ('nchw','channels_last'):
    if f in d: print(f, {k:(v.get('images_per_s'), v.get('ms_per_step'), v.get('error')) for k,v in d[f].items()})
"; done
