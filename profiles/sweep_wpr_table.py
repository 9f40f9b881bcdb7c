"""Tuning aid: time dlmcq_obs_sweep_channel for every warps-per-row choice (DLMCQ_SWEEP_WPR) on weight shapes."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlmc_quant_b200 import _lib  # noqa: E402

h = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
shapes = [(48, 27), (64, 147), (64, 576), (96, 864), (192, 1728), (256, 2304), (512, 4608), (512, 1152), (1280, 1728),
          (2048, 1152), (1000, 2048), (2048, 512), (960, 9), (256, 64), (1024, 256), (8192, 1152), (32768, 1152)]
print(f"{'shape':>16s} " + " ".join(f"wpr={w:<6d}" for w in (1, 2, 4, 8)) + " default")
for c, k in shapes:
    w = torch.randn(c, k, device="cuda") * 0.02
    sc, of = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    row = []
    for wpr in (1, 2, 4, 8, 0):
        if wpr:
            os.environ["DLMCQ_SWEEP_WPR"] = str(wpr)
        else:
            os.environ.pop("DLMCQ_SWEEP_WPR", None)
        run = lambda: _lib.check(h.dlmcq_obs_sweep_channel(w.data_ptr(), c, k, _lib.F32, 4, 1, sc.data_ptr(), of.data_ptr(), st))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            run()
        b.record()
        torch.cuda.synchronize()
        row.append(a.elapsed_time(b) / 20 * 1e3)
    print(f"{str((c, k)):>16s} " + " ".join(f"{t:10.1f}" for t in row))
