"""Standalone kernel sweep (BASELINE.json configs[4] / SURVEY.md 8d "C5"): fake-quant fwd, bwd and the
observers on 2^20 .. 2^30 elements, fp32 and bf16, per-tensor and per-channel layouts, timed with CUDA
events around back-to-back C-ABI launches (pre-built arguments, outputs pre-allocated; tensors below
2^26 elements are L2-resident and flagged).  One JSON line per measurement.

    python profiles/standalone_sweep.py [--max-log2 30] [--quick] > profiles/r01_standalone_sweep.jsonl
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlmc_quant_b200 import _lib  # noqa: E402
from dlmc_quant_b200 import functional as F  # noqa: E402

PEAK = 6550.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
h = _lib.lib()


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def emit(**kw):
    kw["frac_of_measured_peak"] = round(kw["gbps"] / PEAK, 4)
    print(json.dumps(kw), flush=True)


def fq_case(shape, ch_axis, dtype, form, lo, hi, tag):
    n = math.prod(shape)
    es = 4 if dtype == torch.float32 else 2
    iters = max(5, min(200, int(2e9 / (n * es))))
    signed = lo < 0
    x = torch.randn(shape, device="cuda") * (0.02 if signed else 1.5)
    if not signed:
        x = torch.relu(x)
    x = x.to(dtype)
    dy = torch.randn(shape, device="cuda").to(dtype)
    stats = F.obs_stats(x, ch_axis=ch_axis)
    scale, off = F.minmax_from_stats(stats, 4, signed)
    y, dx = torch.empty_like(x), torch.empty_like(x)
    ds = torch.empty(scale.numel(), device="cuda")
    lay = F.layout_of(x, ch_axis)
    qp = _lib.QParams(form, lo, hi, 1 / math.sqrt(n * hi), scale.data_ptr(), off.data_ptr())
    wsn = h.dlmcq_workspace_bytes(C.byref(lay))
    ws = torch.zeros(wsn, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fwd():
        _lib.check(h.dlmcq_fq_forward(x.data_ptr(), y.data_ptr(), None, C.byref(lay), C.byref(qp), st))

    def bwd():
        _lib.check(h.dlmcq_fq_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), ds.data_ptr(), None, C.byref(lay),
                                       C.byref(qp), ws.data_ptr(), wsn, st))

    tf, tb = timed(fwd, iters), timed(bwd, iters)
    common = dict(layout=tag, shape=list(shape), elements=n, dtype=str(dtype).split(".")[-1],
                  l2_resident=bool(n * es * 3 < 100e6))
    emit(op="fq_fwd", us=round(tf * 1e6, 2), gbps=round(2 * es * n / tf / 1e9, 1), gelem_s=round(n / tf / 1e9, 2), **common)
    emit(op="fq_bwd", us=round(tb * 1e6, 2), gbps=round(3 * es * n / tb / 1e9, 1), gelem_s=round(n / tb / 1e9, 2), **common)
    emit(op="fq_fwd+bwd", us=round((tf + tb) * 1e6, 2), gbps=round(5 * es * n / (tf + tb) / 1e9, 1),
         gelem_s=round(n / (tf + tb) / 1e9, 2), **common)


def observer_cases(n, dtype):
    es = 4 if dtype == torch.float32 else 2
    iters = max(5, min(200, int(2e9 / (n * es))))
    x = (torch.relu(torch.randn(n, device="cuda")) * 2).to(dtype)
    lay = F.layout_of(x)
    wsn = h.dlmcq_workspace_bytes(C.byref(lay))
    ws = torch.zeros(wsn, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = torch.empty(4, device="cuda")
    sse = torch.empty(80, device="cuda")
    code = _lib.F32 if dtype == torch.float32 else _lib.BF16
    common = dict(layout="per-tensor", elements=n, dtype=str(dtype).split(".")[-1], l2_resident=bool(n * es < 100e6))

    def stat():
        _lib.check(h.dlmcq_obs_stats(x.data_ptr(), stats.data_ptr(), C.byref(lay), 0, ws.data_ptr(), wsn, st))

    t = timed(stat, iters)
    emit(op="obs_stats(min,max,absmax,abssum)", us=round(t * 1e6, 2), gbps=round(es * n / t / 1e9, 1),
         gelem_s=round(n / t / 1e9, 2), **common)

    def sweep():
        _lib.check(h.dlmcq_obs_sweep_tensor_sse(x.data_ptr(), n, code, stats.data_ptr(), 4, 1, sse.data_ptr(),
                                                ws.data_ptr(), wsn, st))

    t = timed(sweep, max(3, iters // 20))
    emit(op="obs_sweep_tensor(80 candidates)", us=round(t * 1e6, 2), gbps=round(es * n / t / 1e9, 1),
         gelem_s=round(n / t / 1e9, 2), candidate_evals_per_s=round(80 * n / t / 1e12, 3), **common)

    scale = torch.full((1,), 0.3, device="cuda")
    off = torch.zeros(1, device="cuda")
    diff = torch.zeros(1, device="cuda")
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    lay1 = _lib.Layout(1, 1, n, code)
    wsn1 = h.dlmcq_workspace_bytes(C.byref(lay1))
    ws1 = torch.zeros(wsn1, dtype=torch.uint8, device="cuda")

    def l2n():
        flags.zero_()
        _lib.check(h.dlmcq_obs_l2norm_step(x.data_ptr(), 1, n, code, scale.data_ptr(), off.data_ptr(), 0, 15,
                                           diff.data_ptr(), flags.data_ptr(), flags.data_ptr() + 4, ws1.data_ptr(), wsn1, st))

    t = timed(l2n, iters)
    emit(op="obs_l2norm_step(one iteration)", us=round(t * 1e6, 2), gbps=round(es * n / t / 1e9, 1),
         gelem_s=round(n / t / 1e9, 2), **common)


def channel_sweep_case(c, k):
    w = torch.randn(c, k, device="cuda") * 0.02
    scale, off = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        _lib.check(h.dlmcq_obs_sweep_channel(w.data_ptr(), c, k, _lib.F32, 4, 1, scale.data_ptr(), off.data_ptr(), st))

    t = timed(run, 20)
    emit(op="obs_sweep_channel(80 candidates)", layout=f"rows [{c},{k}]", elements=c * k, dtype="float32",
         us=round(t * 1e6, 2), gbps=round(4 * c * k / t / 1e9, 2), gelem_s=round(c * k / t / 1e9, 3),
         candidate_evals_per_s=round(80 * c * k / t / 1e12, 3), l2_resident=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    torch.manual_seed(2333)
    sizes = [20, 22, 24, 26, 28, 30] if not args.quick else [22, 26]
    for lg in [s for s in sizes if s <= args.max_log2]:
        n = 1 << lg
        for dtype in (torch.float32, torch.bfloat16):
            fq_case((n,), None, dtype, _lib.FORM_AFFINE, 0, 15, "per-tensor A4 (QBase form)")
            torch.cuda.empty_cache()
    # per-channel activations [B,C,H,W], ch_axis=1 (QBase 'channel' input types) at ~2^26 elements
    for c, hw in [(64, 56), (256, 28), (2048, 7)] if not args.quick else [(256, 28)]:
        b = max(1, (1 << 26) // (c * hw * hw))
        fq_case((b, c, hw, hw), 1, torch.float32, _lib.FORM_AFFINE, 0, 15, f"per-channel act ch_axis=1 C={c} HW={hw * hw}")
        torch.cuda.empty_cache()
    # per-channel weights [C,K], ch_axis=0 (FSPTQ symmetric form)
    for c, k in [(64, 576), (512, 4608), (2048, 1152), (1000, 2048), (960, 9)] if not args.quick else [(512, 4608)]:
        fq_case((c, k), 0, torch.float32, _lib.FORM_SYM, -7, 7, f"per-channel weight [C={c},K={k}] (FSPTQ form)")
    for lg in ([22, 26, 28] if not args.quick else [24]):
        for dtype in (torch.float32, torch.bfloat16):
            observer_cases(1 << lg, dtype)
            torch.cuda.empty_cache()
    for c, k in [(512, 4608), (2048, 1152), (1280, 1728), (96, 864), (48, 27)]:
        channel_sweep_case(c, k)


if __name__ == "__main__":
    main()
