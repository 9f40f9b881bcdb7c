"""dlmc/quantization/scalar/utils.py with the same names and argument meaning, on the device.

quantize / dequantize / emulate_quantize run the FORM_A1 kernel (utils.py:1-11).  grad_scale,
round_pass and floor_pass keep their autograd behaviour (value from a kernel, straight-through
gradient) - the fused module kernels never call them; they exist for callers of the reference API."""
import torch

from .. import functional as F
from .._lib import FORM_A1, DlmcqError

__all__ = ["quantize", "dequantize", "emulate_quantize", "get_qrange", "grad_scale", "round_pass", "floor_pass",
           "infer_ch_axis"]


def get_qrange(signed, n_bits):
    """utils.py:14-22 (host integers)."""
    if signed:
        max_val = 2 ** (n_bits - 1) - 1
        min_val = -max_val
    else:
        max_val = 2 ** n_bits - 1
        min_val = 0
    return min_val, max_val


def infer_ch_axis(tensor, scale):
    """Which [outer, channels, inner] view reproduces broadcasting `scale` against `tensor`:
    None for a scalar scale, else (first, last) axis of the run of non-1 dims of the scale."""
    if not isinstance(scale, torch.Tensor) or scale.numel() == 1:
        return None
    shape = [1] * (tensor.dim() - scale.dim()) + list(scale.shape)
    axes = [i for i, d in enumerate(shape) if d != 1]
    first, last = axes[0], axes[-1]
    if any(shape[i] != tensor.shape[i] for i in range(first, last + 1)):
        raise DlmcqError(f"scale shape {tuple(scale.shape)} is not a contiguous-run broadcast of {tuple(tensor.shape)}")
    return first, last


def _view3(tensor, scale):
    """tensor reshaped so that the scale's channel run is ONE axis; returns (view, ch_axis)."""
    run = infer_ch_axis(tensor, scale)
    if run is None:
        return tensor, None
    first, last = run
    if first == last:
        return tensor, first
    shp = list(tensor.shape)
    merged = shp[:first] + [int(torch.tensor(shp[first:last + 1]).prod())] + shp[last + 1:]
    return tensor.reshape(merged), first


def quantize(tensor, scale, offset, min_val, max_val):
    """utils.py:1-2 -> integer codes as a float tensor."""
    v, ax = _view3(tensor.contiguous(), scale)
    return F.fq_forward(v, scale, offset, min_val, max_val, FORM_A1, ch_axis=ax, want_codes=True,
                        want_y=False).reshape(tensor.shape)


def dequantize(tensor_q, scale, offset):
    """utils.py:5-6."""
    v, ax = _view3(tensor_q.contiguous(), scale)
    return F.dequantize(v, scale, offset, ch_axis=ax).reshape(tensor_q.shape)


def emulate_quantize(tensor, scale, offset, min_val, max_val):
    """utils.py:9-11."""
    v, ax = _view3(tensor.contiguous(), scale)
    return F.fq_forward(v, scale, offset, min_val, max_val, FORM_A1, ch_axis=ax).reshape(tensor.shape)


class _GradScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return F.grad_scale_value(x.reshape(-1), scale).reshape(x.shape)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


class _StePass(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode):
        return F.ste_value(x, mode)

    @staticmethod
    def backward(ctx, g):
        return g, None


def grad_scale(x, scale):
    """utils.py:24-27: value (x - x*scale) + x*scale, gradient multiplied by `scale`."""
    return _GradScale.apply(x, float(scale))


def round_pass(x):
    """utils.py:29-32: round-half-even value, identity gradient."""
    return _StePass.apply(x, 0)


def floor_pass(x):
    """utils.py:34-37: floor value, identity gradient."""
    return _StePass.apply(x, 1)
