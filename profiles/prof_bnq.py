"""Roofline of the fused BatchNorm(+residual)(+ReLU)->fake-quant op (csrc/bnq_kernels.cu) on ResNet-50's channels-last
activation shapes at batch 128: whole forward (statistics + finalise + apply) and whole backward (reduce + finalise + dx)
through the C ABI, CUDA events, 20 timed calls after 5 warm-ups; algorithmic bytes per element:
    forward : read x twice (statistics, apply) + write a_q = 12 B; + identity read 4 B, + plain output 4 B
    backward: plain chain: read x, d_q twice + write dx = 20 B; residual form: read a, x, d_a, d_q + write dz, then
              read dz, x + write dx = 32 B
Tensors below the 126 MB L2 are reported too (flagged): the cache, not HBM, bounds them."""
import ctypes as C
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlmc_quant_b200 import _lib, functional as F  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.0


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters * 1e-3


def run(shape, resid, plain, dtype=torch.float32):
    n, c, h, w = shape
    rows = n * h * w
    dev = torch.device("cuda")
    es = 4 if dtype is torch.float32 else 2
    mk = lambda: torch.randn(rows, c, device=dev).to(dtype)
    x, idn, a, aq, dq, da, dx, dz = mk(), mk(), mk(), mk(), mk(), mk(), mk(), mk()
    gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    sm, si, dg, db = (torch.empty(c, device=dev) for _ in range(4))
    scale, off, ds = torch.tensor([0.2], device=dev), torch.zeros(1, device=dev), torch.empty(1, device=dev)
    flags = _lib.BNQ_TRAINING | _lib.BNQ_RELU | (_lib.BNQ_RESIDUAL if resid else 0)
    desc = _lib.BnqDesc(rows, c, F._dtype_code(x), flags, 1e-5, 0.1)
    qp = _lib.QParams(_lib.FORM_AFFINE, 0, 15, 1 / math.sqrt(rows * c * 15), scale.data_ptr(), off.data_ptr())
    h_ = _lib.lib()
    nws = h_.dlmcq_bnq_workspace_bytes(C.byref(desc))
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    st = F._stream_ptr()
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None

    def fwd():
        _lib.check(h_.dlmcq_bnq_forward(p(x), p(idn) if resid else None, p(gamma), p(beta), p(rm), p(rv), p(sm), p(si),
                                         p(a) if plain else None, p(aq), C.byref(desc), C.byref(qp), p(ws), nws, st))

    def bwd():
        _lib.check(h_.dlmcq_bnq_backward(p(x), p(a) if resid else None, p(da) if plain else None, p(dq), p(gamma), p(beta),
                                          p(sm), p(si), p(dx), p(dz) if resid else None, p(dg), p(db), p(ds),
                                          C.byref(desc), C.byref(qp), p(ws), nws, st))
    fwd()
    tf, tb = timeit(fwd), timeit(bwd)
    ne = rows * c
    bf = (3 + (1 if resid else 0) + (1 if plain else 0)) * es * ne
    bb = ((8 if plain else 6) if resid else (7 if plain else 5)) * es * ne
    return {"shape": list(shape), "dtype": str(dtype).split(".")[-1], "residual": resid, "plain_out": plain,
            "MB": round(ne * es / 1e6, 1), "l2_resident": ne * es < 100e6,
            "fwd_us": round(tf * 1e6, 1), "fwd_GBps": round(bf / tf / 1e9, 1), "fwd_frac": round(bf / tf / 1e9 / PEAK, 3),
            "bwd_us": round(tb * 1e6, 1), "bwd_GBps": round(bb / tb / 1e9, 1), "bwd_frac": round(bb / tb / 1e9 / PEAK, 3)}


if __name__ == "__main__" and "--one" in sys.argv:
    # for ncu: two launches of every kernel on one HBM-sized plain-chain tensor and one residual-form tensor
    _t = timeit
    timeit = lambda fn, iters=1, warm=1: _t(fn, 1, 1)
    print(json.dumps(run((128, 256, 56, 56), False, False)))
    print(json.dumps(run((128, 256, 56, 56), True, True)))
    print(json.dumps(run((128, 64, 56, 56), False, False)))
    sys.exit(0)

if __name__ == "__main__":
    shapes = [(128, 64, 112, 112), (128, 64, 56, 56), (128, 256, 56, 56), (128, 128, 28, 28), (128, 512, 28, 28),
              (128, 256, 14, 14), (128, 1024, 14, 14), (128, 512, 7, 7), (128, 2048, 7, 7)]
    for s in shapes:
        big = s[1] >= 4 * 64 and s != (128, 256, 14, 14) and s != (128, 512, 7, 7)
        print(json.dumps(run(s, resid=big and s[1] in (256, 512, 1024, 2048), plain=big)), flush=True)
    print(json.dumps(run((128, 256, 56, 56), False, False)), flush=True)
    print(json.dumps(run((128, 256, 56, 56), False, False, torch.bfloat16)), flush=True)
    print(json.dumps(run((128, 256, 56, 56), True, True, torch.bfloat16)), flush=True)
