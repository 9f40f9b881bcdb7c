"""dlmc/quantization/scalar/modules/function.py: the autograd-Function surface, backed by the fused
kernels.  `FakeQuantFunction` is the one the modules use; the Fun* classes keep the reference's
names and `apply` signatures.

Reference defects not reproduced: FunUniformQ.backward unpacks three values from a
two-tuple and returns mask*weight instead of mask*grad (function.py:20,25); FunRootQ.backward returns
one gradient for five inputs (function.py:62).  Both are given the evidently intended straight-through
behaviour here; FunLSQ, the only one whose backward works upstream, is reproduced exactly."""
import torch

from ... import functional as F
from ... import torch_ops  # noqa: F401  (registers torch.ops.dlmcq.* - the path taken under torch.compile)
from ..._lib import FORM_A1, FORM_AFFINE, FORM_SYM, FORM_ZP  # noqa: F401
from ..utils import _view3, infer_ch_axis

__all__ = ["FakeQuantFunction", "fake_quantize", "FunUniformQ", "FunLSQ", "FunRootQ", "FunLQ"]


class FakeQuantFunction(torch.autograd.Function):
    """y = fake_quant(x; scale, offset) with the fused backward: one pass over (x, dy) producing dx and
    the reduced d(scale) (same shape as `scale`).  `form` selects the reference expression."""

    @staticmethod
    def forward(ctx, x, scale, offset, lo, hi, form, g, ch_axis):
        ctx.save_for_backward(x, scale, offset if isinstance(offset, torch.Tensor) else None)
        ctx.cfg = (lo, hi, form, g, ch_axis)
        return F.fq_forward(x, scale, offset, lo, hi, form, g=g, ch_axis=ch_axis)

    @staticmethod
    def backward(ctx, dy):
        x, scale, offset = ctx.saved_tensors
        lo, hi, form, g, ch_axis = ctx.cfg
        dx, ds = F.fq_backward(x, dy, scale, offset, lo, hi, form, g=g, ch_axis=ch_axis)
        if ds.shape != scale.shape:
            ds = ds.reshape(scale.shape)
        return dx, ds if ds.dtype is scale.dtype else ds.to(scale.dtype), None, None, None, None, None, None


def fake_quantize(x, scale, offset, lo, hi, form, g=0.0):
    """Differentiable fake-quant; the channel axis is inferred from the scale's broadcast shape."""
    run = infer_ch_axis(x, scale)
    if torch.compiler.is_compiling() and (run is None or run[0] == run[1]):
        # under torch.compile the quantizer is one opaque registered op (dlmc_quant_b200/torch_ops.py)
        ax = -1 if run is None else run[0]
        off = offset if isinstance(offset, torch.Tensor) else None
        v = x if (x.is_contiguous() or F.dense_as_is(x, None if ax < 0 else ax)) else x.contiguous()
        return torch.ops.dlmcq.fq_forward(v, scale, off, int(lo), int(hi), int(form), float(g), ax)
    if run is None or run[0] == run[1]:                      # the common cases: no reshape, no extra autograd node
        ax = None if run is None else run[0]
        v = x if (x.is_contiguous() or F.dense_as_is(x, ax)) else x.contiguous()   # channels_last passes through
        return FakeQuantFunction.apply(v, scale, offset, lo, hi, form, g, ax)
    v, ax = _view3(x.contiguous(), scale)
    return FakeQuantFunction.apply(v, scale, offset, lo, hi, form, g, ax).reshape(x.shape)


class FunLSQ(torch.autograd.Function):
    """function.py:29-49.  forward: emulate_quantize (A1 form); backward: dw = mid*dy with strict masks
    on w/scale, dscale[1] = g * sum(dy*(lo*below + hi*above + mid*(round(q)-q))), offset ignored."""

    @staticmethod
    def forward(ctx, weight, scale, offset, min_val, max_val, g):
        ctx.save_for_backward(weight, scale)
        ctx.other = g, min_val, max_val
        return F.fq_forward(weight, scale, offset, min_val, max_val, FORM_A1)

    @staticmethod
    def backward(ctx, grad_weight):
        weight, scale = ctx.saved_tensors
        g, min_val, max_val = ctx.other
        dw, ds = F.fq_backward(weight, grad_weight, scale, None, min_val, max_val, FORM_A1, g=g)
        return dw, ds.reshape(1), None, None, None, None


class FunUniformQ(torch.autograd.Function):
    """function.py:9-27: emulate_quantize forward; straight-through backward inside the clamp range."""

    @staticmethod
    def forward(ctx, weight, scale, offset, min_val, max_val):
        ctx.save_for_backward(weight, scale)
        ctx.other = min_val, max_val
        return F.fq_forward(weight, scale, offset, min_val, max_val, FORM_A1)

    @staticmethod
    def backward(ctx, grad_weight):
        weight, scale = ctx.saved_tensors
        min_val, max_val = ctx.other
        dw, _ = F.fq_backward(weight, grad_weight, scale, None, min_val, max_val, FORM_A1, g=0.0)
        return dw, None, None, None, None


class FunRootQ(torch.autograd.Function):
    """function.py:51-62: emulate_quantize forward, identity backward."""

    @staticmethod
    def forward(ctx, weight, scale, offset, min_val, max_val):
        return F.fq_forward(weight, scale, offset, min_val, max_val, FORM_A1)

    @staticmethod
    def backward(ctx, grad_weight):
        return grad_weight, None, None, None, None


class FunLQ(torch.autograd.Function):
    """function.py:64-71: identity both ways."""

    @staticmethod
    def forward(ctx, weight, scale, offset, min_val, max_val, g):
        return weight.view_as(weight)

    @staticmethod
    def backward(ctx, grad_weight):
        return grad_weight, None, None, None, None, None
