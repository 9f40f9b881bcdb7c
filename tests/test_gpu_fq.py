"""GPU parity tests for the fused fake-quant kernels, called through the C ABI (ctypes) and
compared with (a) the golden fixtures minted from the unmodified reference and (b) the oracle
restatement on seeded inputs.  Forward values and integer codes must be BIT-exact; reduced
scale gradients are compared at 1e-5 relative (north_star) with a reduction-order floor."""
import math

import pytest
import torch

from oracle import restate as R
from tests.golden_io import bits_equal, first_mismatch, load

pytestmark = pytest.mark.gpu

A1, AFFINE, ZP, SYM = 0, 1, 2, 3


def F():
    from dlmc_quant_b200 import functional
    return functional


def exact(a, b, what=""):
    assert bits_equal(a.cpu(), b.cpu()), f"{what}: {first_mismatch(a.cpu(), b.cpu())}"


def red_close(mine, ref, abs_sum=None, rtol=1e-5):
    """|mine-ref| <= rtol*|ref| + 4e-7*abs_sum: the second term is the fp32 reduction-order floor
    (abs_sum = sum of |terms|); without it a nearly cancelling sum has no meaningful relative error."""
    mine, ref = mine.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    tol = rtol * ref.abs()
    if abs_sum is not None:
        tol = tol + 4e-7 * abs_sum.detach().double().cpu().reshape(-1)
    bad = ~((mine - ref).abs() <= tol) & ~(torch.isnan(mine) & torch.isnan(ref))
    assert not bad.any(), f"reduced value mismatch: mine={mine[bad][:4]}, ref={ref[bad][:4]}, tol={tol[bad][:4]}"


def dev(t):
    return t.cuda() if isinstance(t, torch.Tensor) else t


def floor_of(dy, qabs, mul=1.0, rows=None):
    """Reduction-order floor for a scale gradient: the reference forms sum(dy*code) and sum(dy*u)
    separately in fp32 and subtracts, so its own error scales with sum|dy|*max|code|."""
    a = dy.abs()
    a = a.reshape(rows, -1).sum(1) if rows else a.sum().reshape(1)
    return a * float(qabs) * float(mul)


# --------------------------------------------------------------------------------------
# golden fixtures (reference's own outputs)
UTILS = load("utils")


@pytest.mark.parametrize("name", sorted(n for n in UTILS if n.startswith("a1_")))
def test_golden_a1(name):
    c = UTILS[name]
    lo, hi = c.meta["lo"], c.meta["hi"]
    x = c.inp["x"]
    ax = 0 if name.endswith("_pc") else None
    y, codes = F().fq_forward(dev(x), dev(c.inp["scale"]), dev(c.inp["offset"]), lo, hi, A1, ch_axis=ax, want_codes=True)
    exact(codes, c.out["codes"], "codes")
    exact(y, c.out["y"], "y")
    exact(F().dequantize(codes, dev(c.inp["scale"]), dev(c.inp["offset"]), ch_axis=ax), c.out["y"], "dequantize")


QBASE = load("qbase")


@pytest.mark.parametrize("name", sorted(QBASE))
def test_golden_qbase_quantizers(name):
    c = QBASE[name]
    q = c.meta["qconfig"]
    x, w = c.inp["x"], c.inp["weight"]
    ilo, ihi = R.qrange(q["input"]["args"]["signed"], q["input"]["args"]["n_bits"])
    wlo, whi = R.qrange(q["weight"]["args"]["signed"], q["weight"]["args"]["n_bits"])
    g_i, g_w = R.lsq_g(x.numel(), ihi), R.lsq_g(w.numel(), whi)
    wax = 0 if c.meta.get("per_channel_weight") else None
    qx = F().fq_forward(dev(x), dev(c.out["param_in_scale"]), dev(c.out["buf_in_offset"]), ilo, ihi, AFFINE, g=g_i)
    qw = F().fq_forward(dev(w), dev(c.out["param_wt_scale"]), dev(c.out["buf_wt_offset"]), wlo, whi, AFFINE, g=g_w,
                        ch_axis=wax)
    exact(qx, c.out["qx"], "qx")
    exact(qw, c.out["qw"], "qw")
    dx, ds = F().fq_backward(dev(x), dev(c.out["d_qx"]), dev(c.out["param_in_scale"]), dev(c.out["buf_in_offset"]),
                             ilo, ihi, AFFINE, g=g_i)
    assert torch.allclose(dx.cpu(), c.out["dx"], rtol=1e-6, atol=0)
    assert torch.equal(dx.cpu() == 0, c.out["dx"] == 0)
    red_close(ds, c.out["grad_in_scale"], abs_sum=floor_of(c.out["d_qx"], ihi, g_i), rtol=1e-5)
    dw, dsw = F().fq_backward(dev(w), dev(c.out["d_qw"]), dev(c.out["param_wt_scale"]), dev(c.out["buf_wt_offset"]),
                              wlo, whi, AFFINE, g=g_w, ch_axis=wax)
    assert torch.allclose(dw.cpu(), c.out["grad_weight"], rtol=1e-6, atol=0)
    red_close(dsw, c.out["grad_wt_scale"], abs_sum=floor_of(c.out["d_qw"], whi, g_w, rows=w.shape[0] if wax == 0 else None),
              rtol=1e-5)


FUNLSQ = load("funlsq")


@pytest.mark.parametrize("name", sorted(FUNLSQ))
def test_golden_funlsq(name):
    c = FUNLSQ[name]
    lo, hi, g = c.meta["lo"], c.meta["hi"], c.meta["g"]
    y = F().fq_forward(dev(c.inp["w"]), dev(c.inp["scale"]), dev(c.inp["offset"]), lo, hi, A1)
    exact(y, c.out["y"], "y")
    dw, ds = F().fq_backward(dev(c.inp["w"]), dev(c.inp["dy"]), dev(c.inp["scale"]), dev(c.inp["offset"]), lo, hi, A1, g=g)
    exact(dw, c.out["dw"], "dw")
    red_close(ds, c.out["dscale"], abs_sum=floor_of(c.inp["dy"], max(abs(lo), hi), g), rtol=1e-5)


FSPTQ = load("fsptq")


@pytest.mark.parametrize("name", sorted(n for n in FSPTQ if "ada" not in n))
def test_golden_fsptq_quantizers(name):
    c = FSPTQ[name]
    q = c.meta["qconfig"]
    x, w = c.inp["x"], c.inp["weight"]
    ilo, ihi = R.qrange(q["input"]["args"]["signed"], q["input"]["args"]["n_bits"])
    wlo, whi = R.qrange(q["weight"]["args"]["signed"], q["weight"]["args"]["n_bits"])
    s_in, o_in, s_w = c.out["param_in_scale"], c.out["buf_in_offset"], c.out["param_wt_scale"]
    exact(F().fq_forward(dev(x), dev(s_in), dev(o_in), ilo, ihi, ZP), c.out["qx"], "qx")
    exact(F().fq_forward(dev(w), dev(s_w), None, wlo, whi, SYM, ch_axis=0), c.out["qw"], "qw")
    dx, ds = F().fq_backward(dev(x), dev(c.out["d_qx"]), dev(s_in), dev(o_in), ilo, ihi, ZP)
    assert torch.allclose(dx.cpu(), c.out["dx"], rtol=1e-6, atol=0)
    red_close(ds, c.out["grad_in_scale"], abs_sum=floor_of(c.out["d_qx"], ihi), rtol=1e-5)
    dw, dsw = F().fq_backward(dev(w), dev(c.out["d_qw"]), dev(s_w), None, wlo, whi, SYM, ch_axis=0)
    assert torch.allclose(dw.cpu(), c.out["grad_weight"], rtol=1e-6, atol=0)
    red_close(dsw, c.out["grad_wt_scale"], abs_sum=floor_of(c.out["d_qw"], whi, rows=w.shape[0]), rtol=1e-5)


@pytest.mark.parametrize("name", sorted(n for n in FSPTQ if "ada" in n))
def test_golden_adaround(name):
    c = FSPTQ[name]
    q = c.meta["qconfig"]
    w = c.inp["weight"]
    wlo, whi = R.qrange(True, q["weight"]["args"]["n_bits"])
    s_w = c.out["param_wt_scale"]
    alpha = F().adaround_init_alpha(dev(w), dev(s_w))
    ref_alpha = R.adaround_init_alpha(w, s_w)
    assert torch.allclose(alpha.cpu(), ref_alpha, rtol=2e-6, atol=1e-6, equal_nan=True)
    a = dev(ref_alpha)
    y = F().adaround_forward(dev(w), a, dev(s_w), wlo, whi, soft=True)
    assert torch.allclose(y.cpu(), c.out["qw"], rtol=1e-6, atol=1e-9)      # sigmoid: transcendental, not bit-exact
    if "qw_eval" in c.out:
        exact(F().adaround_forward(dev(w), a, dev(s_w), wlo, whi, soft=False), c.out["qw_eval"], "hard rounding")
    dalpha, ds = F().adaround_backward(dev(w), a, dev(c.out["d_qw"]), dev(s_w), wlo, whi)
    red_close(ds, c.out["grad_wt_scale"], abs_sum=floor_of(c.out["d_qw"], whi, rows=w.shape[0]), rtol=1e-5)
    if "grad_alpha" in c.out:
        assert torch.allclose(dalpha.cpu(), c.out["grad_alpha"], rtol=1e-5, atol=1e-9)


# --------------------------------------------------------------------------------------
# oracle on seeded inputs: forms x dtypes x layouts, ragged / unaligned / empty
def make_qparams(form, x, ch_axis, signed, bits, gen):
    lo, hi = R.qrange(signed, bits)
    if ch_axis is None:
        shape = [1]
    else:
        shape = [1] * x.dim()
        shape[ch_axis] = x.shape[ch_axis]
    n = math.prod(shape)
    base = x.abs().max().item() / max(hi, 1) if x.numel() else 1.0
    scale = (torch.rand(n, generator=gen) * 0.8 + 0.4).reshape(shape) * max(base, 1e-3)
    if form == ZP:
        off = torch.randint(0, 5, (n,), generator=gen).float().reshape(shape)
    elif form == SYM:
        off = None
    else:
        off = (torch.randn(n, generator=gen) * 0.1).reshape(shape)
    return lo, hi, scale, off


def oracle_fwd(form, x, scale, off, lo, hi, g):
    if form == A1:
        return R.codes_a1(x, scale, off, lo, hi), R.emulate_a1(x, scale, off, lo, hi)
    if form == AFFINE:
        return R.fq_affine_codes(x, scale, off, lo, hi, g), R.fq_affine(x, scale, off, lo, hi, g)
    if form == ZP:
        return R.fq_zp_codes(x, scale, off, lo, hi), R.fq_zp(x, scale, off, lo, hi)
    return R.fq_sym_codes(x, scale, lo, hi), R.fq_sym(x, scale, lo, hi)


def oracle_bwd(form, x, scale, off, lo, hi, g, dy):
    xs = x.detach().clone().requires_grad_(True)
    ss = scale.detach().clone().requires_grad_(True)
    if form == AFFINE:
        y = R.fq_affine(xs, ss, off, lo, hi, g)
    elif form == ZP:
        y = R.fq_zp(xs, ss, off, lo, hi)
    else:
        y = R.fq_sym(xs, ss, lo, hi)
    dx, ds = torch.autograd.grad(y, (xs, ss), dy)
    return dx, ds


SHAPES = [
    ((1 << 20) + 3, None),       # per-tensor, ragged tail
    ((7,), None),                # smaller than one vector
    ((64, 147), 0),              # conv1 rows (odd length -> unaligned rows)
    ((48, 27), 0),               # RepVGG stem rows
    ((32, 9), 0),                # depthwise rows
    ((16, 4608), 0),             # longest ResNet row
    ((3, 5000), 0),              # row longer than one segment
    ((4, 8, 7, 7), 1),           # activations, channel axis 1, 49-element planes (scalar channel-major kernel)
    ((5, 6, 8, 8), 1),           # 64-element planes: whole 128-bit vectors (vectorised channel-major kernel)
    ((3, 4, 14, 14), 1),         # 196-element planes: vectorised for fp32, scalar for bf16 (196 % 8 != 0)
    ((2, 6, 56, 56), 1),         # long planes: tiled per-channel kernels (rows span several 16 KB tiles)
    ((3, 5, 16, 16), 1),         # 64-vector planes: 16-17 rows per tile (fp32 tiled; bf16 warp-per-row)
    ((3, 7, 20, 20), 1),         # 100-vector planes, ragged last tile
]


@pytest.mark.parametrize("form", [A1, AFFINE, ZP, SYM])
@pytest.mark.parametrize("shape,ch_axis", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_vs_oracle(form, shape, ch_axis, dtype):
    gen = torch.Generator().manual_seed(2333 + form)
    shape = (shape,) if isinstance(shape, int) else shape
    signed = form in (A1, SYM)
    x = torch.randn(shape, generator=gen) * (0.05 if signed else 1.5)
    if not signed:
        x = torch.relu(x)
    x = x.to(dtype)
    lo, hi, scale, off = make_qparams(form, x.float(), ch_axis, signed, 4, gen)
    g = R.lsq_g(max(x.numel(), 1), hi)
    codes_ref, y_ref = oracle_fwd(form, x.float(), scale, off, lo, hi, g)
    y, codes = F().fq_forward(dev(x), dev(scale), dev(off), lo, hi, form, g=g, ch_axis=ch_axis, want_codes=True)
    assert y.dtype == dtype
    exact(codes.float(), codes_ref.to(dtype).float(), "codes")
    exact(y.float(), y_ref.to(dtype).float(), "y")          # bf16: one RNE rounding of the fp32 result


@pytest.mark.parametrize("form", [AFFINE, ZP, SYM])
@pytest.mark.parametrize("shape,ch_axis", SHAPES)
def test_backward_vs_oracle(form, shape, ch_axis):
    gen = torch.Generator().manual_seed(4666 + form)
    shape = (shape,) if isinstance(shape, int) else shape
    signed = form == SYM
    x = torch.randn(shape, generator=gen) * (0.05 if signed else 1.5)
    if not signed:
        x = torch.relu(x)
    dy = torch.randn(shape, generator=gen)
    lo, hi, scale, off = make_qparams(form, x, ch_axis, signed, 4, gen)
    g = R.lsq_g(x.numel(), hi)
    dx_ref, ds_ref = oracle_bwd(form, x, scale, off, lo, hi, g, dy)
    dx, ds = F().fq_backward(dev(x), dev(dy), dev(scale), dev(off), lo, hi, form, g=g, ch_axis=ch_axis)
    assert torch.equal(dx.cpu() == 0, dx_ref == 0), "clamp mask differs"
    assert torch.allclose(dx.cpu(), dx_ref, rtol=1e-6, atol=0)      # reference computes (dy*s)/s: 1-ulp wobble
    # reduction-order floor: sum of |dy| * max|code| per channel
    if ch_axis is None:
        abs_sum = (dy.abs().sum() * max(abs(lo), abs(hi))).reshape(1)
    else:
        red = [d for d in range(x.dim()) if d != ch_axis]
        abs_sum = dy.abs().sum(dim=red) * max(abs(lo), abs(hi))
    if form == AFFINE:
        abs_sum = abs_sum * g
    red_close(ds, ds_ref, abs_sum=abs_sum)
    # and against the exact (fp64) sum of the fp32 per-element terms: the fused kernel itself must be
    # accurate to 1e-5 relative with only a tiny floor (it never forms the two cancelling sums)
    if form == AFFINE:
        sp = R.grad_scale(scale, g)
        u = (x - off) / sp
        code = R.round_ste(u.clamp(lo, hi))
        inside = ((u >= lo) & (u <= hi))
        terms = dy.double() * (code.double() - torch.where(inside, u, torch.zeros_like(u)).double()) * g
    else:
        v = x / scale
        t = R.round_ste(v) + (off if form == ZP else 0)
        inside = (t >= lo) & (t <= hi)
        deq = t.clamp(lo, hi) - (off if form == ZP else 0)
        terms = dy.double() * (deq.double() - torch.where(inside, v, torch.zeros_like(v)).double())
    if ch_axis is None:
        exact_sum, tabs = terms.sum().reshape(1), terms.abs().sum().reshape(1)
    else:
        red = [d for d in range(x.dim()) if d != ch_axis]
        exact_sum, tabs = terms.sum(dim=red), terms.abs().sum(dim=red)
    red_close(ds, exact_sum, abs_sum=tabs * 0.5)


def test_backward_bf16_matches_fp32_math():
    gen = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn(1 << 18, generator=gen) * 1.5).bfloat16()
    dy = torch.randn(1 << 18, generator=gen).bfloat16()
    scale, off = torch.tensor([0.31]), torch.tensor([0.0])
    g = R.lsq_g(x.numel(), 15)
    dx_ref, ds_ref = oracle_bwd(AFFINE, x.float(), scale, off, 0, 15, g, dy.float())
    dx, ds = F().fq_backward(dev(x), dev(dy), dev(scale), dev(off), 0, 15, AFFINE, g=g)
    assert dx.dtype == torch.bfloat16
    exact(dx.float(), dx_ref.bfloat16().float(), "dx bf16")
    red_close(ds, ds_ref, abs_sum=(dy.float().abs().sum() * 15 * g).reshape(1))


def test_unaligned_views_and_empty():
    gen = torch.Generator().manual_seed(11)
    base = torch.relu(torch.randn(4099, generator=gen)).cuda()
    x = base[1:]                                   # 4-byte aligned only
    scale, off = torch.tensor([0.2]), torch.tensor([0.05])
    y = F().fq_forward(x, dev(scale), dev(off), 0, 15, AFFINE, g=0.01)
    exact(y, R.fq_affine(x.cpu(), scale, off, 0, 15, 0.01), "unaligned fwd")
    e = torch.empty(0, device="cuda")
    assert F().fq_forward(e, dev(scale), dev(off), 0, 15, AFFINE).numel() == 0
    dx, ds = F().fq_backward(e, e, dev(scale), dev(off), 0, 15, AFFINE, g=0.1)
    assert dx.numel() == 0 and float(ds) == 0.0


def test_special_values_all_forms():
    """NaN / inf / signed zero / ties / denormals / zero scale behave like the eager chain."""
    from oracle.make_golden import special_values
    v = torch.cat([special_values(), special_values() * 0.37, torch.zeros(5)])
    for form, lo, hi in [(A1, -7, 7), (AFFINE, 0, 15), (ZP, 0, 15), (SYM, -7, 7), (AFFINE, -127, 127)]:
        for s, o in [(1.0, 0.0), (0.37, 0.25), (0.0, 0.0), (1e-30, 0.0)]:
            scale, off = torch.tensor([s]), torch.tensor([o if form != ZP else float(round(o * 12))])
            codes_ref, y_ref = oracle_fwd(form, v, scale, off, lo, hi, 0.003)
            y, codes = F().fq_forward(dev(v), dev(scale), dev(off if form != SYM else None), lo, hi, form, g=0.003,
                                      want_codes=True)
            exact(codes, codes_ref, f"codes form={form} s={s} o={o}")
            exact(y, y_ref, f"y form={form} s={s} o={o}")


def test_workspace_is_left_reusable():
    """Two backward calls in a row on one stream give identical, deterministic results."""
    gen = torch.Generator().manual_seed(5)
    x = dev(torch.relu(torch.randn(1 << 21, generator=gen)))
    dy = dev(torch.randn(1 << 21, generator=gen))
    s, o = dev(torch.tensor([0.2])), dev(torch.tensor([0.0]))
    r = [F().fq_backward(x, dy, s, o, 0, 15, AFFINE, g=0.001) for _ in range(3)]
    assert all(bits_equal(r[0][1].cpu(), k[1].cpu()) for k in r[1:])
    assert all(torch.equal(r[0][0], k[0]) for k in r[1:])


@pytest.mark.parametrize("shape", [(3, 6, 32, 32), (4, 8, 8, 8), (3, 5, 7, 7), (2, 4, 16, 16)])
def test_backward_bf16_per_channel_layouts(shape):
    """bf16 per-channel activations through every kernel family (tiled: 32x32 planes = 128 vectors; vectorised
    channel-major: 8x8; scalar channel-major: 7x7; warp-per-row: 16x16 = 32 vectors): fp32 math on the up-converted
    values, one rounding of dx, fp32 scale gradients."""
    gen = torch.Generator().manual_seed(17)
    x = (torch.relu(torch.randn(shape, generator=gen)) * 1.5).bfloat16()
    dy = torch.randn(shape, generator=gen).bfloat16()
    scale, off = R.obs_minmax_channel(x.float(), 4, False, ch_axis=1)
    g = R.lsq_g(x.numel(), 15)
    dx_ref, ds_ref = oracle_bwd(AFFINE, x.float(), scale, off, 0, 15, g, dy.float())
    _, y_ref = oracle_fwd(AFFINE, x.float(), scale, off, 0, 15, g)
    y = F().fq_forward(dev(x), dev(scale), dev(off), 0, 15, AFFINE, g=g, ch_axis=1)
    exact(y.float(), y_ref.bfloat16().float(), "y bf16")
    dx, ds = F().fq_backward(dev(x), dev(dy), dev(scale), dev(off), 0, 15, AFFINE, g=g, ch_axis=1)
    assert dx.dtype == torch.bfloat16
    exact(dx.float(), dx_ref.bfloat16().float(), "dx bf16")
    red_close(ds, ds_ref.reshape(-1), abs_sum=dy.float().abs().sum(dim=(0, 2, 3)) * 15 * g)


# --------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE sizes; oracle too slow on CPU here)
@pytest.mark.parametrize("n", [1 << 26])
def test_full_size_properties(n):
    torch.manual_seed(2333)
    x = torch.relu(torch.randn(n, device="cuda")) * 2
    dy = torch.randn(n, device="cuda")
    stats = F().obs_stats(x)
    scale, off = F().minmax_from_stats(stats, 4, False)
    g = R.lsq_g(n, 15)
    y, codes = F().fq_forward(x, scale, off, 0, 15, AFFINE, g=g, want_codes=True)
    # codes are integers in range; dequantised output reproduces them
    assert bool((codes == codes.round()).all()) and float(codes.min()) >= 0 and float(codes.max()) <= 15
    # idempotence: fake-quantising the output again changes nothing
    assert torch.equal(F().fq_forward(y, scale, off, 0, 15, AFFINE, g=g), y)
    # the same chain evaluated by eager torch on the GPU (the "GPU reference") agrees bit for bit
    y_eager = R.fq_affine(x, scale, off, 0, 15, g)
    assert torch.equal(y_eager, y)
    # backward: linear in dy, mask identical to eager autograd
    dx, ds = F().fq_backward(x, dy, scale, off, 0, 15, AFFINE, g=g)
    dx2, ds2 = F().fq_backward(x, dy * 2, scale, off, 0, 15, AFFINE, g=g)
    assert torch.equal(dx2, dx * 2)
    assert abs(float(ds2) - 2 * float(ds)) <= 1e-5 * abs(float(ds2)) + 1e-6
    xs = x.clone().requires_grad_(True)
    ss = scale.clone().requires_grad_(True)
    dxe, dse = torch.autograd.grad(R.fq_affine(xs, ss, off, 0, 15, g), (xs, ss), dy)
    assert torch.equal(dxe == 0, dx == 0)
    red_close(ds, dse, abs_sum=(dy.abs().sum() * 15 * g).reshape(1))


def test_fast_division_is_ieee_exact():
    """6.4e9 (x, s) pairs per domain: the fast path's 3-instruction division == __fdiv_rn bit for bit."""
    import ctypes as C
    from dlmc_quant_b200 import _lib
    h = _lib.lib()
    for narrow in (1, 0):
        bad = torch.zeros(1, dtype=torch.int64, device="cuda")
        for rep in range(4):
            _lib.check(h.dlmcq_selftest_fastdiv(2333 + rep, 148 * 8, 1775, narrow, C.c_void_p(bad.data_ptr()), None))
        torch.cuda.synchronize()
        assert int(bad) == 0, f"{int(bad)} quotients differ from IEEE division (narrow={narrow})"


def test_deferred_finalize_equals_inline():
    """backward_partials + finalize_many == backward, bit for bit (same partials, same summation order)."""
    from dlmc_quant_b200.functional import DeferredScaleGrads
    gen = torch.Generator().manual_seed(21)
    sizes = [1 << 20, 12345, (1 << 22) + 7, 64]
    dq = DeferredScaleGrads("cuda", len(sizes))
    outs, refs = [], []
    for i, n in enumerate(sizes):
        x = dev(torch.relu(torch.randn(n, generator=gen)) * 2)
        dy = dev(torch.randn(n, generator=gen))
        s, o = dev(torch.tensor([0.21 + 0.01 * i])), dev(torch.tensor([0.0]))
        g = R.lsq_g(n, 15)
        dx = torch.empty_like(x)
        ds = torch.zeros(1, device="cuda")
        dq.backward(i, x, dy, dx, s, o, 0, 15, AFFINE, g)
        outs.append((dx, ds))
        refs.append(F().fq_backward(x, dy, s, o, 0, 15, AFFINE, g=g))
    dq.finalize([ds for _, ds in outs])
    for (dx, ds), (rdx, rds) in zip(outs, refs):
        assert torch.equal(dx, rdx)
        exact(ds, rds, "deferred dscale")


@pytest.mark.parametrize("form,lo,hi,shape,ax", [(AFFINE, 0, 15, (3, 5, 7, 7), None), (SYM, -7, 7, (16, 27), 0),
                                                  (ZP, 0, 255, (1001,), None), (A1, -127, 127, (8, 33), 0),
                                                  (AFFINE, 0, 15, (2, 6, 5, 5), 1)])
def test_integer_export_round_trip(form, lo, hi, shape, ax):
    """packed int4 / int8 codes == the fp32-valued codes of the forward; import(export(x)) == fq_forward(x)."""
    gen = torch.Generator().manual_seed(31 + form)
    signed = lo < 0
    x = torch.randn(shape, generator=gen) * (0.05 if signed else 1.5)
    if not signed:
        x = torch.relu(x)
    _, _, scale, off = make_qparams(form, x, ax, signed, 4 if hi <= 15 else 8, gen)
    g = R.lsq_g(x.numel(), hi)
    y, codes = F().fq_forward(dev(x), dev(scale), dev(off), lo, hi, form, g=g, ch_axis=ax, want_codes=True)
    for pack4 in ([False, True] if hi - lo <= 15 else [False]):
        packed = F().export_codes(dev(x), dev(scale), dev(off), lo, hi, form, g=g, ch_axis=ax, pack4=pack4)
        if pack4:
            b = packed.cpu().to(torch.int32)
            nib = torch.stack([b & 15, b >> 4], 1).reshape(-1)[: x.numel()]
            ints = torch.where(nib > 7, nib - 16, nib) if signed else nib
            assert packed.numel() == (x.numel() + 1) // 2
        else:
            ints = packed.cpu().to(torch.int32).reshape(-1)
            assert packed.dtype == (torch.int8 if signed else torch.uint8)
        assert torch.equal(ints, codes.cpu().reshape(-1).to(torch.int32)), "integer codes differ"
        back = F().import_codes(packed, tuple(shape), dev(scale), dev(off), lo, hi, form, g=g, ch_axis=ax, pack4=pack4)
        exact(back, y, "import(export(x)) vs fq_forward")
