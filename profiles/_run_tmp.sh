timeout 600 python -m pytest tests/test_gpu_rootq_obs.py -q -x -k "kth or percentile or resident" 2>&1 | tail -8 | cut -c1-300
timeout 300 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from dlmc_quant_b200 import functional as F
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
for n in (1<<24, 1<<26, 1<<28):
    x=torch.relu(torch.randn(n, device='cuda'))*2
    k_hi=int(0.9999*n); k_lo=n+1-k_hi
    a=t(lambda:F.kth_values(x,[k_lo,k_hi],fast=True)); b=t(lambda:F.kth_values(x,[k_lo,k_hi],fast=False))
    a1=t(lambda:F.kth_values(x,[k_hi],abs_input=True,fast=True))
    print(f"n=2^{n.bit_length()-1} post-ReLU fp32: one-read path {a:.0f} us ({4*n/a/1e6:.2f} TB/s of one read), three-pass {b:.0f} us; one rank |x|: {a1:.0f} us")
PY
