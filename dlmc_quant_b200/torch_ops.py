"""`torch.library` registration of the hot-path ops (namespace `dlmcq`) over the same C ABI.

SURVEY.md 8b, last row: the Python side registers the kernels as custom ops "so torch.compile / CUDA graphs can capture
them".  Eager code keeps the direct autograd.Function path (functional.py: it is the cheaper one per call); while
TorchDynamo is tracing, `fake_quantize` (scalar/modules/function.py) routes through these ops instead: opaque to the
compiler, with shape-only (fake / meta) implementations and explicit autograd formulas, so a quantised model
compiles without graph breaks at the quantizers.

    torch.ops.dlmcq.fq_forward(x, scale, offset, lo, hi, form, g, ch_axis) -> y
    torch.ops.dlmcq.fq_backward(x, dy, scale, offset, lo, hi, form, g, ch_axis) -> (dx, dscale)
(ch_axis = -1: per-tensor.)"""
from typing import Optional, Tuple

import torch

from . import functional as F

__all__ = ["fq_forward", "fq_backward"]


def _ax(ch_axis):
    return None if ch_axis < 0 else ch_axis


@torch.library.custom_op("dlmcq::fq_forward", mutates_args=())
def fq_forward(x: torch.Tensor, scale: torch.Tensor, offset: Optional[torch.Tensor], lo: int, hi: int, form: int,
               g: float, ch_axis: int) -> torch.Tensor:
    return F.fq_forward(x, scale, offset, lo, hi, form, g=g, ch_axis=_ax(ch_axis))


@fq_forward.register_fake
def _(x, scale, offset, lo, hi, form, g, ch_axis):
    return torch.empty_like(x)


@torch.library.custom_op("dlmcq::fq_backward", mutates_args=())
def fq_backward(x: torch.Tensor, dy: torch.Tensor, scale: torch.Tensor, offset: Optional[torch.Tensor], lo: int, hi: int,
                form: int, g: float, ch_axis: int) -> Tuple[torch.Tensor, torch.Tensor]:
    dx, ds = F.fq_backward(x, dy, scale, offset, lo, hi, form, g=g, ch_axis=_ax(ch_axis))
    return dx, ds.reshape(scale.shape).to(scale.dtype)


@fq_backward.register_fake
def _(x, dy, scale, offset, lo, hi, form, g, ch_axis):
    return torch.empty_like(x), torch.empty_like(scale)


def _setup(ctx, inputs, output):
    x, scale, offset, lo, hi, form, g, ch_axis = inputs
    ctx.save_for_backward(x, scale, offset)
    ctx.cfg = (lo, hi, form, g, ch_axis)


def _backward(ctx, dy):
    x, scale, offset = ctx.saved_tensors
    lo, hi, form, g, ch_axis = ctx.cfg
    dx, ds = torch.ops.dlmcq.fq_backward(x, dy, scale, offset, lo, hi, form, g, ch_axis)
    return dx, ds, None, None, None, None, None, None


fq_forward.register_autograd(_backward, setup_context=_setup)
