"""Timing of the exact k-th-value selection (percentile observer) on 2^26 elements; torch.kthvalue beside it."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlmc_quant_b200 import functional as F  # noqa: E402

n = 1 << 26
for name, x in (("relu(randn)*2", torch.relu(torch.randn(n, device="cuda")) * 2), ("randn", torch.randn(n, device="cuda")),
                ("bf16 randn", torch.randn(n, device="cuda").bfloat16())):
    ranks = [n // 10000, n - n // 10000]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = {}
    for fast in (True, False):
        for _ in range(2):
            F.kth_values(x, ranks, fast=fast)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            v = F.kth_values(x, ranks, fast=fast)
        b.record()
        torch.cuda.synchronize()
        times[fast] = a.elapsed_time(b) / 10
    t = times[False]
    es = x.element_size()
    print(f"kth_values {name:14s} two ranks, ONE-READ path (sample bracket + collect + cooperative resolve): "
          f"{times[True] * 1e3:7.1f} us = {es * n / times[True] / 1e6:6.0f} GB/s of one read")
    a.record()
    w = torch.stack([x.float().kthvalue(k)[0] for k in ranks])
    b.record()
    torch.cuda.synchronize()
    print(f"kth_values {name:14s} two ranks: {t * 1e3:7.1f} us = {3 * es * n / t / 1e6:6.0f} GB/s over three reads; "
          f"equal to torch.kthvalue: {torch.equal(v, w)} (torch: {a.elapsed_time(b) * 1e3:.0f} us)")
