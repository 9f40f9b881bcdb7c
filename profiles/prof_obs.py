"""Profiling driver for the kernels that are not on the flat streaming path: the 80-candidate sweeps,
the l2norm iteration, the statistics pass and the per-channel (rows / channel-major) fake-quant kernels.
Run plain first (prints CUDA-event times), then under ncu (profiles/README.md).

    python profiles/prof_obs.py [reps]
"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlmc_quant_b200 import _lib  # noqa: E402
from dlmc_quant_b200 import functional as F  # noqa: E402

h = _lib.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(2333)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(name, fn, bytes_, extra=""):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3 / reps
    print(f"{name:44s} {t * 1e6:9.1f} us  {bytes_ / t / 1e9:8.1f} GB/s {extra}")
    return t


# ---- per-tensor observers on 2^26 fp32 elements -------------------------------------------------
n = 1 << 26
x = torch.relu(torch.randn(n, device="cuda")) * 2
lay = F.layout_of(x)
wsn = h.dlmcq_workspace_bytes(C.byref(lay))
ws = torch.zeros(wsn, dtype=torch.uint8, device="cuda")
stats = torch.empty(4, device="cuda")
sse = torch.empty(80, device="cuda")
timed("obs_stats flat 2^26", lambda: _lib.check(h.dlmcq_obs_stats(x.data_ptr(), stats.data_ptr(), C.byref(lay), 0,
                                                                    ws.data_ptr(), wsn, st)), 4 * n)
t = timed("obs_sweep_tensor 2^26", lambda: _lib.check(h.dlmcq_obs_sweep_tensor_sse(
    x.data_ptr(), n, _lib.F32, stats.data_ptr(), 8, 1, sse.data_ptr(), ws.data_ptr(), wsn, st)), 4 * n)
print(f"    candidate-evals/s: {80 * n / t / 1e12:.3f} T")
scale = torch.full((1,), 0.3, device="cuda")
off = torch.zeros(1, device="cuda")
diff = torch.zeros(1, device="cuda")
flags = torch.zeros(2, dtype=torch.int32, device="cuda")
lay1 = _lib.Layout(1, 1, n, _lib.F32)
wsn1 = h.dlmcq_workspace_bytes(C.byref(lay1))
ws1 = torch.zeros(wsn1, dtype=torch.uint8, device="cuda")


def l2n():
    flags.zero_()
    _lib.check(h.dlmcq_obs_l2norm_step(x.data_ptr(), 1, n, _lib.F32, scale.data_ptr(), off.data_ptr(), 0, 15,
                                       diff.data_ptr(), flags.data_ptr(), flags.data_ptr() + 4, ws1.data_ptr(), wsn1, st))


timed("obs_l2norm_step flat 2^26", l2n, 4 * n)

# ---- per-channel sweeps on weight-shaped rows ---------------------------------------------------
for c, k in [(512, 4608), (2048, 1152), (1280, 1728), (96, 864)]:
    w = torch.randn(c, k, device="cuda") * 0.02
    sc, of = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    t = timed(f"obs_sweep_channel [{c},{k}]", lambda: _lib.check(h.dlmcq_obs_sweep_channel(
        w.data_ptr(), c, k, _lib.F32, 4, 1, sc.data_ptr(), of.data_ptr(), st)), 4 * c * k)
    print(f"    candidate-evals/s: {80 * c * k / t / 1e12:.3f} T")


# ---- per-channel activations --------------------------------------------------------------------
def fq_case(shape, ch_axis, tag):
    m = math.prod(shape)
    xx = torch.relu(torch.randn(shape, device="cuda") * 1.5)
    dy = torch.randn(shape, device="cuda")
    s4 = F.obs_stats(xx, ch_axis=ch_axis)
    sc, of = F.minmax_from_stats(s4, 4, False)
    y, dx = torch.empty_like(xx), torch.empty_like(xx)
    ds = torch.empty(sc.numel(), device="cuda")
    ly = F.layout_of(xx, ch_axis)
    qp = _lib.QParams(_lib.FORM_AFFINE, 0, 15, 1 / math.sqrt(m * 15), sc.data_ptr(), of.data_ptr())
    wn = h.dlmcq_workspace_bytes(C.byref(ly))
    w_ = torch.zeros(wn, dtype=torch.uint8, device="cuda")
    timed(f"fq_fwd {tag}", lambda: _lib.check(h.dlmcq_fq_forward(xx.data_ptr(), y.data_ptr(), None, C.byref(ly),
                                                                   C.byref(qp), st)), 8 * m)
    timed(f"fq_bwd {tag}", lambda: _lib.check(h.dlmcq_fq_backward(xx.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                                                                    ds.data_ptr(), None, C.byref(ly), C.byref(qp),
                                                                    w_.data_ptr(), wn, st)), 12 * m)
    timed(f"obs_stats {tag}", lambda: _lib.check(h.dlmcq_obs_stats(xx.data_ptr(), s4.data_ptr(), C.byref(ly), 0,
                                                                     w_.data_ptr(), wn, st)), 4 * m)


for c, hw in [(64, 56), (256, 28), (2048, 7)]:
    b = max(1, (1 << 26) // (c * hw * hw))
    fq_case((b, c, hw, hw), 1, f"per-channel act C={c} HW={hw * hw}")
    torch.cuda.empty_cache()
