"""CPU: the two oracles against each other on randomised inputs.  oracle/restate.py (eager torch ops + autograd, the
reference's own way of computing) and oracle/fq_oracle.c (scalar C loops, closed-form gradients) are independent
statements of the same path; beyond the fixed fixtures they must agree on arbitrary shapes, bit-widths, scales and
special values - forward bit for bit, reduced gradients to reduction-order accuracy."""
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import restate as R
from tests.golden_io import bits_equal

SPECIAL = [0.0, -0.0, float("nan"), float("inf"), float("-inf"), 1e-45, -1e-45, 3.4e38, 0.5, 1.5, 2.5, -0.5, -1.5, 7.5, 8.5]


def _case(seed):
    gen = torch.Generator().manual_seed(seed)
    r = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=gen))
    channels, inner, outer = r(1, 6), r(1, 70), r(1, 3)
    bits, signed = r(2, 8), bool(r(0, 1))
    x = torch.randn(outer, channels, inner, generator=gen) * (10 ** (r(-3, 1)))
    if not signed:
        x = x.abs()
    flat = x.reshape(-1)
    idx = torch.randperm(flat.numel(), generator=gen)[:min(len(SPECIAL), flat.numel())]
    if seed % 3 == 0:
        flat[idx] = torch.tensor(SPECIAL[:idx.numel()])
    scale = (torch.rand(channels, generator=gen) + 0.05) * float(x[torch.isfinite(x)].abs().max().clamp(min=1e-3)) / (2 ** bits)
    offset = torch.randn(channels, generator=gen) * 0.1 if not signed else torch.zeros(channels)
    dy = torch.randn(x.shape, generator=gen)
    lo, hi = R.qrange(signed, bits)
    return x, dy, scale, offset, lo, hi, channels, inner


def _bc(v, x):
    return v.reshape(1, -1, 1).expand_as(x) if v.numel() > 1 else v


@pytest.mark.parametrize("seed", range(40))
def test_forward_forms_agree_bit_for_bit(seed):
    x, _, scale, offset, lo, hi, channels, inner = _case(seed)
    g = R.lsq_g(x.numel(), max(abs(lo), abs(hi)))
    sb, ob = _bc(scale, x), _bc(offset, x)
    zp = torch.round(offset * 20)
    want = {0: (R.codes_a1(x, sb, ob, lo, hi), R.emulate_a1(x, sb, ob, lo, hi)),
            1: (R.fq_affine_codes(x, sb, ob, lo, hi, g), R.fq_affine(x, sb, ob, lo, hi, g)),
            2: (R.fq_zp_codes(x, sb, _bc(zp, x), lo, hi), R.fq_zp(x, sb, _bc(zp, x), lo, hi)),
            3: (R.fq_sym_codes(x, sb, lo, hi), R.fq_sym(x, sb, lo, hi))}
    for form, (codes_t, y_t) in want.items():
        off = {0: offset, 1: offset, 2: zp, 3: None}[form]
        y, codes = CO.fq_forward(x.numpy(), scale.numpy(), None if off is None else off.numpy(), form, lo, hi, g,
                                 channels=channels, inner=inner)
        assert bits_equal(torch.from_numpy(codes).reshape(-1), codes_t.detach().reshape(-1)), (seed, form, "codes")
        assert bits_equal(torch.from_numpy(y).reshape(-1), y_t.detach().reshape(-1)), (seed, form, "y")


@pytest.mark.parametrize("seed", range(1, 40, 3))          # seeds without injected NaN / inf
def test_closed_form_gradients_agree_with_autograd(seed):
    x, dy, scale, offset, lo, hi, channels, inner = _case(seed)
    qmax = max(abs(lo), abs(hi))
    g = R.lsq_g(x.numel(), qmax)
    floor = 4e-7 * dy.abs().sum(dim=(0, 2)) * qmax
    # AFFINE (QBase): d scale carries the grad_scale factor g
    xs, ss = x.clone().requires_grad_(True), scale.clone().requires_grad_(True)
    yt = R.fq_affine(xs, ss.reshape(1, -1, 1), offset.reshape(1, -1, 1), lo, hi, g)
    dx_t, ds_t = torch.autograd.grad(yt, (xs, ss), dy)
    dx, ds = CO.fq_backward(x.numpy(), dy.numpy(), scale.numpy(), offset.numpy(), 1, lo, hi, g, channels, inner)
    assert np.array_equal(dx.reshape(-1) == 0, dx_t.reshape(-1).numpy() == 0)
    assert np.allclose(dx.reshape(-1), dx_t.reshape(-1).numpy(), rtol=1e-6, atol=0)
    assert np.all(np.abs(ds - ds_t.double().numpy()) <= 2e-5 * np.abs(ds_t.double().numpy()) + (floor * g).numpy() + 1e-12), seed
    # SYM (FSPTQ weights)
    xs, ss = x.clone().requires_grad_(True), scale.clone().requires_grad_(True)
    yt = R.fq_sym(xs, ss.reshape(1, -1, 1), lo, hi)
    dx_t, ds_t = torch.autograd.grad(yt, (xs, ss), dy)
    dx, ds = CO.fq_backward(x.numpy(), dy.numpy(), scale.numpy(), None, 3, lo, hi, 0.0, channels, inner)
    assert np.array_equal(dx.reshape(-1) == 0, dx_t.reshape(-1).numpy() == 0)
    assert np.all(np.abs(ds - ds_t.double().numpy()) <= 2e-5 * np.abs(ds_t.double().numpy()) + floor.numpy() + 1e-12), seed


@pytest.mark.parametrize("seed", range(8))
def test_percentile_oracles_agree(seed):
    """The extension observer's two CPU statements: torch.kthvalue (oracle/restate.py) and a plain-C sort."""
    gen = torch.Generator().manual_seed(100 + seed)
    n = int(torch.randint(1, 5000, (1,), generator=gen))
    x = torch.randn(n, generator=gen) * 3
    if seed % 2:
        x = torch.relu(x)
    for k in sorted({1, n, (n + 1) // 2, max(1, n - n // 100)}):
        assert CO.kth_value(x.numpy(), k) == float(x.kthvalue(k)[0])
        assert CO.kth_value(x.numpy(), k, abs_input=True) == float(x.abs().kthvalue(k)[0])
    signed = bool(seed % 2 == 0)
    s, o = R.obs_percentile_tensor(x if signed else x.abs(), 8, signed, 99.0)
    k_hi = min(n, max(1, math.ceil(0.99 * n)))
    v = x if signed else x.abs()
    if signed:
        assert float(s) == np.float32(CO.kth_value(v.numpy(), k_hi, abs_input=True)) / np.float32(127)
    else:
        lo_v, hi_v = CO.kth_value(v.numpy(), n + 1 - k_hi), CO.kth_value(v.numpy(), k_hi)
        assert float(s) == (np.float32(hi_v) - np.float32(lo_v)) / np.float32(255) and float(o) == lo_v
