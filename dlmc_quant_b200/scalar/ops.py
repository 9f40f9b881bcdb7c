"""dlmc/quantization/scalar/ops.py: the PTQ observers, name-dispatched exactly like the reference
(`get_qparams_tensor(tensor, qtype, **kwargs)` -> `quantize_<qtype>`), computed by the observer kernels.

Differences that are deliberate and documented in DESIGN.md:
  * results stay on the tensor's device (the reference returns `torch.tensor(0)` on the CPU as the
    signed offset, ops.py:24) - no host round trip, same values;
  * when torch.distributed is initialised, statistics are all-reduced so that every rank derives the
    same qparams (the reference initialises each DDP rank from its local batch);
  * l2norm loops are bounded (`max_iters`): the reference's `while diff > eps` never terminates on
    some inputs (the unsigned per-channel form oscillates forever on data with negative values);
  * `quantize_l2norm_pixel` is not provided: in the reference it raises NameError on every input (ops.py:237
    calls `emulate_quantize`, which ops.py never imports), its best-scale bookkeeping is dead code
    (ops.py:242-244 assigns best_mse to itself) and nothing in the repo selects it.
"""
import math

import torch

from .. import dist as qdist
from .. import functional as F
from .utils import get_qrange, quantize

__all__ = ["get_qparams_tensor", "get_qparams_output", "quantize_minmax_tensor", "quantize_minmax_channel",
           "quantize_minmax_pixel", "quantize_l2loss_tensor", "quantize_l2loss_channel", "quantize_l2norm_tensor",
           "quantize_l2norm_channel", "quantize_l2norm_output", "quantize_l2norm_output_channel",
           "quantize_percentile_tensor"]


def get_qparams_output(input, weight, module, qtype, **kwargs):
    """ops.py:11-13."""
    return globals()[f"quantize_{qtype}"](input, weight, module, **kwargs)


def get_qparams_tensor(tensor, qtype, **kwargs):
    """ops.py:15-18."""
    return globals()[f"quantize_{qtype}"](tensor, **kwargs)


def _process_channel(tensor, ch_axis):
    """ops.py:112-118: channel-major rows (a copy when ch_axis != 0) and the broadcast shape."""
    new_shape = [1] * tensor.dim()
    new_shape[ch_axis] = -1
    return tensor.transpose(0, ch_axis).reshape(tensor.shape[ch_axis], -1), new_shape


def _stats(tensor, ch_axis=None, abs_input=False):
    return qdist.sync_stats(F.obs_stats(tensor, ch_axis=ch_axis, abs_input=abs_input))


def _check_nonneg(stats):
    # ops.py:29,45,132,161 `assert (min_val >= 0).all()` - one host sync, calibration time only
    assert bool((stats[:, 0] >= 0).all()), "allow_offset=False requires non-negative data"


def quantize_minmax_tensor(tensor, n_bits, signed, allow_offset=True):
    """ops.py:20-34 -> 0-dim (scale, offset)."""
    stats = _stats(tensor)
    if not signed and not allow_offset:
        _check_nonneg(stats)
    scale, offset = F.minmax_from_stats(stats, n_bits, signed, allow_offset)
    return scale.reshape(()), offset.reshape(())


def quantize_minmax_channel(tensor, n_bits, signed, ch_axis=0, allow_offset=True):
    """ops.py:121-140 -> (scale, offset) shaped [1,..,C,..,1].  No transpose copy: the statistics
    kernel reduces the [outer, C, inner] view in place."""
    stats = _stats(tensor, ch_axis=ch_axis)
    if not signed and not allow_offset:
        _check_nonneg(stats)
    scale, offset = F.minmax_from_stats(stats, n_bits, signed, allow_offset)
    new_shape = [1] * tensor.dim()
    new_shape[ch_axis] = -1
    return scale.reshape(new_shape), offset.reshape(new_shape)


def quantize_minmax_pixel(tensor, n_bits, signed, allow_offset=True):
    """ops.py:142-167: one (scale, offset) per kernel position, reduced over (Cout, Cin).  The unsigned
    branch takes |x| before min/max - a reference quirk, reproduced."""
    new_shape = [tensor.shape[2], tensor.shape[3]] if tensor.dim() == 4 else [tensor.shape[2]]
    rows = tensor.reshape(tensor.shape[0] * tensor.shape[1], -1).t().contiguous()      # [P, Cout*Cin]
    stats = _stats(rows, ch_axis=0, abs_input=not signed)
    if not signed and not allow_offset:
        _check_nonneg(stats)
    scale, offset = F.minmax_from_stats(stats, n_bits, signed, allow_offset)
    return scale.reshape(new_shape), offset.reshape(new_shape)


def percentile_ranks(numel, percentile):
    """1-based ranks of the upper / lower clipping points: k_hi = ceil(p/100 * N), k_lo = N + 1 - k_hi."""
    k_hi = min(numel, max(1, math.ceil(percentile / 100.0 * numel)))
    return numel + 1 - k_hi, k_hi


def quantize_percentile_tensor(tensor, n_bits, signed, percentile=99.99, allow_offset=True):
    """Percentile-clipping observer - an EXTENSION named by the project's north star; dlmc's ops.py has no such
    function (dispatchable as qtype "percentile_tensor" like the others).  Exact order statistics (3-pass radix
    select, == torch.kthvalue) instead of min/max, then the min/max formulas of ops.py:20-34:
      signed:   scale = kth(|x|, k_hi) / (2^(n-1) - 1), offset = 0
      unsigned: scale = (kth(x, k_hi) - lo) / (2^n - 1), offset = lo, lo = kth(x, k_lo) (0 if not allow_offset).
    Under torch.distributed the per-pass histograms are all-reduced: the statistics of the union of all ranks."""
    n = tensor.numel()
    reduce_hist = None
    w = qdist.world_size()
    if w > 1:
        import torch.distributed as dist
        cnt = torch.tensor([n], dtype=torch.int64, device=tensor.device)
        dist.all_reduce(cnt)
        n = int(cnt)
        reduce_hist = lambda h: dist.all_reduce(h)
    k_lo, k_hi = percentile_ranks(n, percentile)
    if signed:
        a = F.kth_values(tensor, [k_hi], abs_input=True, reduce_hist=reduce_hist)
        stats = torch.stack([a[0], a[0], a[0], a[0]]).reshape(1, 4)
    else:
        v = F.kth_values(tensor, [k_lo, k_hi], reduce_hist=reduce_hist)
        stats = torch.stack([v[0], v[1], v[1].abs(), v[1]]).reshape(1, 4)
        if not allow_offset:
            _check_nonneg(stats)
    scale, offset = F.minmax_from_stats(stats.contiguous(), n_bits, signed, allow_offset)
    return scale.reshape(()), offset.reshape(())


def quantize_l2loss_tensor(tensor, n_bits, signed, allow_offset=True):
    """ops.py:36-68: signed -> the min/max answer; unsigned -> 80-candidate clip sweep in one read."""
    if signed:
        return quantize_minmax_tensor(tensor, n_bits, True)
    if not allow_offset:
        _check_nonneg(_stats(tensor))
    scale, offset, _ = F.sweep_tensor(tensor, n_bits, allow_offset, reduce_stats=qdist.sync_stats,
                                      reduce_sse=qdist.sync_sse)
    return scale.reshape(()), offset.reshape(())


def quantize_l2loss_channel(tensor, n_bits, signed, ch_axis=0):
    """ops.py:169-196: per-channel sweep, rows staged once in shared memory.  Under data parallelism the
    weights are replicated, so the rows are split over the ranks (ceil(C / world) each) and the per-channel
    (scale, offset) pairs all-gathered - each row is swept once per job, not once per rank."""
    rows, new_shape = _process_channel(tensor, ch_axis)
    if ch_axis == 0:        # replicated weights: shard the rows over the ranks and all-gather the qparams
        scale, offset = qdist.rows_sharded(rows, lambda blk: F.sweep_channel(blk, n_bits, signed, rows.shape[0]))
    else:                   # activations differ per rank: local (the reference's per-rank behaviour)
        scale, offset = F.sweep_channel(rows, n_bits, signed)
    return scale.reshape(new_shape), offset.reshape(new_shape)


def quantize_l2norm_tensor(tensor, n_bits, signed, max_iters=1000):
    """ops.py:71-83."""
    scale, offset = quantize_minmax_tensor(tensor, n_bits, signed, allow_offset=True)
    lo, hi = get_qrange(signed, n_bits)
    scale, _, _ = F.l2norm_fixed_point(tensor.reshape(1, -1), scale.reshape(1), offset.reshape(1), lo, hi, max_iters)
    return scale.reshape(()), offset


def quantize_l2norm_channel(tensor, n_bits, signed, ch_axis=0, max_iters=1000):
    """ops.py:198-215."""
    rows, new_shape = _process_channel(tensor, ch_axis)
    scale, offset = quantize_minmax_channel(rows, n_bits, signed, ch_axis=0, allow_offset=True)
    lo, hi = get_qrange(signed, n_bits)
    scale, _, _ = F.l2norm_fixed_point(rows, scale.reshape(-1), offset.reshape(-1), lo, hi, max_iters)
    return scale.reshape(new_shape), offset.reshape(new_shape)


def _l2_loss(a, b):
    """trainer/loss/loss.py:22-24."""
    return ((a - b) ** 2).sum(axis=1).mean()


def quantize_l2norm_output(input, weight, module, n_bits, signed, patience=1000):
    """ops.py:85-109: the error is measured on module._forward_func(input, q(w)); a convolution per
    iteration dominates, so the loop stays on the host and only q(w) runs in our kernels."""
    output = module._forward_func(input, weight)
    scale, offset = quantize_minmax_tensor(weight, n_bits, signed, allow_offset=True)
    lo, hi = get_qrange(signed, n_bits)
    diff, best_mse, best_scale, count = float("inf"), float("inf"), scale, 0
    while diff > 1e-5:
        if count == patience:
            break
        weight_q = quantize(weight, scale, offset, lo, hi)
        output_q = module._forward_func(input, weight_q)
        mse = _l2_loss(output, output_q)
        new_scale = (output_q * output).mean(axis=0).sum() / (output_q * output_q + 1e-7).mean(axis=0).sum()
        diff = float((new_scale - scale).abs() / scale)
        scale = new_scale
        if mse < best_mse:
            best_mse, best_scale = mse, scale
        count += 1
    return best_scale, offset


def quantize_l2norm_output_channel(input, weight, module, n_bits, signed, ch_axis=0, patience=1000):
    """ops.py:252-292."""
    new_shape = [1] * weight.dim()
    new_shape[ch_axis] = -1
    output = module._forward_func(input, weight)
    batch, channel = output.shape[0], output.shape[1]
    output = output.reshape(batch, channel, -1)
    scale, offset = quantize_minmax_channel(weight, n_bits, signed, ch_axis=ch_axis, allow_offset=True)
    lo, hi = get_qrange(signed, n_bits)
    diff, best_mse, best_scale, count = float("inf"), float("inf"), scale, 0
    while diff > 1e-5:
        if count == patience:
            break
        weight_q = quantize(weight, scale, offset, lo, hi)
        output_q = module._forward_func(input, weight_q).reshape(batch, channel, -1)
        new_scale = ((output * output_q).sum(axis=(0, 2)) / (output_q * output_q + 1e-7).sum(axis=(0, 2))).reshape(scale.shape)
        mse = _l2_loss(output, output_q)
        diff = float(((new_scale - scale) ** 2).sum().sqrt() / (scale ** 2).sum().sqrt())
        if mse < best_mse:
            best_mse, best_scale = mse, scale
        scale = new_scale
        count += 1
    return best_scale.reshape(new_shape), offset
