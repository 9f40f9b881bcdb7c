"""dlmc/quantization/scalar/RootQ/function.py: the RootQ helper surface.

The RootQ modules call the fused kernels (RootQActFunction / RootQWeightFunction below).  The
reference's free functions (`clipping`, `torch_phi_function`, `sgn`, `dequantize`, `RoundWithGradient`,
`phi_function`) are kept under the same names as small composites for callers of that API; the
modules never route through them."""
import torch
import torch.nn.functional as TF

from ... import functional as F

__all__ = ["RootQActFunction", "RootQWeightFunction", "RoundWithGradient", "clipping", "torch_phi_function",
           "phi_function", "sgn", "dequantize"]


class RootQActFunction(torch.autograd.Function):
    """RootQ/base.py:108-111 with the closed-form backward of SURVEY.md A.5:
    dx = dy*1[x>=0]*1[x<=upper];  d in_scale = m*g*sum dy*(I - xq/s + Q*1[x>upper])."""

    @staticmethod
    def forward(ctx, x, in_scale, state):
        ctx.save_for_backward(x, state)
        return F.rootq_act_forward(x, state)

    @staticmethod
    def backward(ctx, dy):
        x, state = ctx.saved_tensors
        dx, grads = F.rootq_act_backward(x, dy, state)
        return dx, grads.reshape(()), None


class RootQWeightFunction(torch.autograd.Function):
    """RootQ/base.py:146-155: clipping, floor_pass interval, root-function estimator, sgn, dequantize -
    forward is the nearest grid point; backward is dw = dy*dc/dw*(1 + delta/2*dp/dz) plus the three
    reduced gradients d wt_upper, d wt_lower, d wt_alpha in the same pass."""

    @staticmethod
    def forward(ctx, w, upper, lower, alpha, state):
        ctx.save_for_backward(w, state)
        return F.rootq_wt_forward(w, state)

    @staticmethod
    def backward(ctx, dy):
        w, state = ctx.saved_tensors
        dw, grads = F.rootq_wt_backward(w, dy, state)
        return dw, grads[0].reshape(()), grads[1].reshape(()), grads[2].reshape(()), None


class RoundWithGradient(torch.autograd.Function):
    """function.py:5-12: value sgn(x), gradient identity."""

    @staticmethod
    def forward(ctx, x):
        return F.ste_value(x, 2)

    @staticmethod
    def backward(ctx, g):
        return g


def sgn(x):
    """function.py:58-61."""
    return RoundWithGradient.apply(x)


def clipping(x, upper, lower):
    """function.py:15-20 (relu trick, not clamp)."""
    x = x + TF.relu(lower - x)
    x = x - TF.relu(x - upper)
    return x


def torch_phi_function(x, mi, alpha, delta):
    """function.py:22-32."""
    alpha = alpha + TF.relu(1e-4 - alpha)
    alpha = alpha - TF.relu(alpha - 1)
    x = x - mi
    s = x / (torch.abs(x) + 1e-5)
    k = 2 / delta
    return torch.pow(k * abs(x) + 1e-5, alpha) * s


class phi_function(torch.autograd.Function):
    """function.py:35-56 (the unused alternative estimator), kept for API completeness."""

    @staticmethod
    def forward(ctx, x, mi, alpha, delta):
        ctx.save_for_backward(x, mi, alpha, delta)
        x = 2 * (x - mi) / delta
        deltax = torch.max(x) - torch.min(x) + 1e-6
        x = (x / deltax + 0.5)
        return x.round() * 2 - 1

    @staticmethod
    def backward(ctx, g):
        x, mi, alpha, delta = ctx.saved_tensors
        x = 2 * (x - mi) / delta
        s = x / (abs(x) + 1e-6)
        grad_x = ((abs(x) + 1e-2) ** (alpha - 1)) * alpha * 2 / delta * g
        grad_alpha = torch.log(abs(x) + 1e-2) * ((abs(x) + 1e-2) ** alpha) * s * g
        grad_delta = -1 * grad_x * x
        return grad_x, None, grad_alpha, grad_delta


def dequantize(x, lower_bound, delta, interval):
    """function.py:63-67."""
    return ((x + 1) / 2 + interval) * delta + lower_bound
