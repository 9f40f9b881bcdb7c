timeout 600 python -m pytest tests/test_gpu_rootq_obs.py -q -x 2>&1 | tail -8
timeout 300 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from dlmc_quant_b200 import functional as F
torch.manual_seed(0)
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
for shape in [(512,4608),(2048,512),(64,576),(1,1<<20)]:
    w=(torch.randn(shape)*0.05).cuda()
    s0,o0=F.minmax_from_stats(F.obs_stats(w,ch_axis=0),4,True)
    r=F.l2norm_fixed_point(w,s0,o0,-7,7,resident=True); s=F.l2norm_fixed_point(w,s0,o0,-7,7,resident=False)
    print(shape,'iters',r[1],s[1],'resident %.0f us  stepwise %.0f us'%(t(lambda:F.l2norm_fixed_point(w,s0,o0,-7,7,resident=True)), t(lambda:F.l2norm_fixed_point(w,s0,o0,-7,7,resident=False))))
import bench
rows=[(torch.randn(w[0], int(torch.tensor(w[1:]).prod()))*0.05).cuda() for _,_,w in bench.resnet50_layers()]
print('R50 54 tensors: grouped sweep %.0f us, per-tensor %.0f us'%(t(lambda:F.sweep_channel_grouped(rows,4,True)), t(lambda:[F.sweep_channel(r,4,True) for r in rows])))
PY
