python -m pytest tests/test_gpu_bnq.py -q -x -k "not rewired" 2>&1 | tail -3
for v in "" "DLMCQ_BNQ_DXV=1" "DLMCQ_BNQ_U=2" "DLMCQ_BNQ_TILES=1" "DLMCQ_BNQ_TILES=8"; do echo "== $v"; env $v python profiles/prof_bnq.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['shape'], d['dtype'][:4], 'res' if d['residual'] else 'pla', 'fwd', d['fwd_us'], d['fwd_frac'], 'bwd', d['bwd_us'], d['bwd_frac'])
"; done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_bnq_launches_c.csv python profiles/prof_bnq.py --one > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bnq_bwd_dx" -c 4 -o gpurun_out/r02_bnq_dx2 python profiles/prof_bnq.py --one > gpurun_out/ncu_full.log 2>&1
