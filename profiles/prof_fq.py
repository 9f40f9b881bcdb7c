"""Profiling driver: a few launches of the streaming kernels on 2^26 fp32 elements (256 MB per tensor,
larger than L2).  Run plain first, then under ncu (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlmc_quant_b200 import functional as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(2333)
x = torch.relu(torch.randn(n, device="cuda")) * 2
dy = torch.randn(n, device="cuda")
stats = F.obs_stats(x)
scale, off = F.minmax_from_stats(stats, 4, False)
g = 1.0 / (n * 15) ** 0.5
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for r in range(reps):
    ev[0].record()
    y = F.fq_forward(x, scale, off, 0, 15, F.FORM_AFFINE, g=g)
    ev[1].record()
    dx, ds = F.fq_backward(x, dy, scale, off, 0, 15, F.FORM_AFFINE, g=g)
    ev[2].record()
    st = F.obs_stats(x)
    ev[3].record()
    torch.cuda.synchronize()
    f, b, s = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    print(f"rep {r}: fwd {f*1e3:.1f} us = {8*n/f/1e6:.0f} GB/s | bwd {b*1e3:.1f} us = {12*n/b/1e6:.0f} GB/s | "
          f"stats {s*1e3:.1f} us = {4*n/s/1e6:.0f} GB/s")
# plain copy for comparison (the MEASURED_PEAKS denominator is a copy)
c = torch.empty_like(x)
for r in range(3):
    ev[0].record(); c.copy_(x); ev[1].record(); torch.cuda.synchronize()
    t = ev[0].elapsed_time(ev[1])
    print(f"copy_: {t*1e3:.1f} us = {8*n/t/1e6:.0f} GB/s")

# the reference's eager op chain (oracle restatement of modules/base.py:96-102 + autograd) on the same GPU:
# the "PyTorch eager on B200" comparator of SURVEY.md 8d.  Test infrastructure, not a product path.
from oracle import restate as R  # noqa: E402
xs, ss = x.clone().requires_grad_(True), scale.clone().requires_grad_(True)
for r in range(3):
    ev[0].record()
    ye = R.fq_affine(xs, ss, off, 0, 15, g)
    ev[1].record()
    ye.backward(dy)
    ev[2].record()
    torch.cuda.synchronize()
    f, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"eager chain rep {r}: fwd {f*1e3:.1f} us = {8*n/f/1e6:.0f} GB/s | bwd {b*1e3:.1f} us = {12*n/b/1e6:.0f} GB/s (algorithmic bytes)")
    xs.grad = ss.grad = None
