"""Tuning aid: per-channel activation fake-quant / statistics for small planes, channel-major kernel vs warp-per-row
kernel (DLMCQ_CMAJ_MAX_INNER=1 disables the channel-major path).   python profiles/cmaj_threshold.py"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlmc_quant_b200 import _lib  # noqa: E402
from dlmc_quant_b200 import functional as F  # noqa: E402

h = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
tag = "rows kernel" if os.environ.get("DLMCQ_CMAJ_MAX_INNER") == "1" else "channel-major"
for c, hw in [(2048, 7), (1024, 14), (512, 14), (256, 15), (1024, 8), (512, 12)]:
    b = max(2, (1 << 26) // (c * hw * hw))
    x = torch.relu(torch.randn(b, c, hw, hw, device="cuda") * 1.5)
    dy = torch.randn_like(x)
    s4 = F.obs_stats(x, ch_axis=1)
    sc, of = F.minmax_from_stats(s4, 4, False)
    y, dx, ds = torch.empty_like(x), torch.empty_like(x), torch.empty(c, device="cuda")
    lay = F.layout_of(x, 1)
    qp = _lib.QParams(_lib.FORM_AFFINE, 0, 15, 1 / math.sqrt(x.numel() * 15), sc.data_ptr(), of.data_ptr())
    wn = h.dlmcq_workspace_bytes(C.byref(lay))
    ws = torch.zeros(wn, dtype=torch.uint8, device="cuda")
    fns = {"fwd": (lambda: h.dlmcq_fq_forward(x.data_ptr(), y.data_ptr(), None, C.byref(lay), C.byref(qp), st), 8),
           "bwd": (lambda: h.dlmcq_fq_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), ds.data_ptr(), None, C.byref(lay),
                                               C.byref(qp), ws.data_ptr(), wn, st), 12),
           "stats": (lambda: h.dlmcq_obs_stats(x.data_ptr(), s4.data_ptr(), C.byref(lay), 0, ws.data_ptr(), wn, st), 4)}
    out = []
    for name, (fn, bpe) in fns.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(e) / 10 * 1e-3
        out.append(f"{name} {bpe * x.numel() / t / 1e9:6.0f} GB/s")
    print(f"{tag:14s} [B={b},C={c},{hw}x{hw}]  " + "  ".join(out))
