"""CPU check of the algebra behind the integer-code layer product (oracle/restate.py::code_gemm, the definition
dlmcq_qgemm is held to bit for bit on the GPU): for every activation form the factored expression
alpha[n] * (integer dot product) + beta[n] equals the reference's product of the two fake-quantised tensors
(modules/base.py:140 `_forward_func(q_input, q_weight)`) to 1e-5 relative of sum_k |y_a * y_w|."""
import pytest
import torch

from oracle import restate as R

A1, AFFINE, ZP, SYM = 0, 1, 2, 3


def _setup(a_form, w_form, n_bits, a_signed, per_channel, seed):
    g = torch.Generator().manual_seed(seed)
    m, n, k = 37, 19, 80
    x = torch.randn(m, k, generator=g)
    if not a_signed:
        x = torch.relu(x) + 0.05 * torch.rand(m, k, generator=g)
    w = torch.randn(n, k, generator=g) * 0.05
    a_lo, a_hi = R.qrange(a_signed, n_bits)
    w_lo, w_hi = R.qrange(True, n_bits)
    s_a = (x.abs().max() / a_hi * 0.9).reshape(1)
    off = {ZP: torch.tensor([3.0]), A1: torch.tensor([0.03]), AFFINE: torch.tensor([0.03]), SYM: torch.zeros(1)}[a_form]
    s_w = (w.abs().amax(dim=1, keepdim=True) / w_hi + 1e-6) if per_channel else (w.abs().max() / w_hi).reshape(1)
    g_a = R.lsq_g(x.numel(), a_hi) if a_form == AFFINE else 0.0
    g_w = R.lsq_g(w.numel(), w_hi) if w_form == AFFINE else 0.0
    if a_form == AFFINE:
        qa, ca, m_a = R.fq_affine(x, s_a, off, a_lo, a_hi, g_a), R.fq_affine_codes(x, s_a, off, a_lo, a_hi, g_a), R.grad_scale(s_a, g_a)
    elif a_form == ZP:
        qa, ca, m_a = R.fq_zp(x, s_a, off, a_lo, a_hi), R.fq_zp_codes(x, s_a, off, a_lo, a_hi), s_a
    elif a_form == A1:
        qa, ca, m_a = R.emulate_a1(x, s_a, off, a_lo, a_hi), R.codes_a1(x, s_a, off, a_lo, a_hi), s_a
    else:
        qa, ca, m_a = R.fq_sym(x, s_a, a_lo, a_hi), R.fq_sym_codes(x, s_a, a_lo, a_hi), s_a
    if w_form == AFFINE:
        qw, cw, m_w = (R.fq_affine(w, s_w, torch.zeros(1), w_lo, w_hi, g_w), R.fq_affine_codes(w, s_w, torch.zeros(1), w_lo, w_hi, g_w),
                       R.grad_scale(s_w, g_w))
    else:
        qw, cw, m_w = R.fq_sym(w, s_w, w_lo, w_hi), R.fq_sym_codes(w, s_w, w_lo, w_hi), s_w
    o_a = off if a_form in (A1, AFFINE) else torch.zeros(1)
    z_a = off if a_form == ZP else torch.zeros(1)
    bias = torch.randn(n, generator=g)
    return qa, qw, ca, cw, m_a, o_a, z_a, m_w.reshape(-1), bias


@pytest.mark.parametrize("a_form,w_form", [(AFFINE, AFFINE), (ZP, SYM), (A1, SYM), (SYM, SYM), (AFFINE, SYM)])
@pytest.mark.parametrize("n_bits", [4, 8])
@pytest.mark.parametrize("per_channel", [True, False])
def test_factored_product_equals_reference_product(a_form, w_form, n_bits, per_channel):
    for a_signed in ((False, True) if a_form == SYM else (False,)):
        qa, qw, ca, cw, m_a, o_a, z_a, m_w, bias = _setup(a_form, w_form, n_bits, a_signed, per_channel, seed=n_bits)
        got = R.code_gemm(ca, cw, m_a, o_a, z_a, m_w, bias)
        ref = R.layer_product_reference(qa, qw, bias)
        mag = qa.double().abs() @ qw.double().abs().t() + bias.double().abs()
        err = ((got.double() - ref).abs() / mag).max().item()
        assert err <= 1e-5, err
        # and the reference's own fp32 evaluation is no closer to the exact value than the factored form
        ref32 = R.layer_product_reference(qa, qw, bias, dtype=torch.float32)
        err32 = ((ref32.double() - ref).abs() / mag).max().item()
        assert err <= max(4 * err32, 2e-7)
    assert got.dtype == torch.float32


def test_code_gemm_relu_and_integer_exactness():
    g = torch.Generator().manual_seed(0)
    ca = torch.randint(0, 256, (9, 4096), generator=g)
    cw = torch.randint(-127, 128, (5, 4096), generator=g)
    out = R.code_gemm(ca, cw, 1.0, 0.0, 0.0, torch.ones(5))
    assert torch.equal(out, (ca @ cw.t()).float())            # |acc| up to 1.3e8: one RN conversion, nothing else
    assert (R.code_gemm(ca, cw, 1.0, 0.0, 0.0, torch.ones(5), relu=True) >= 0).all()


@pytest.mark.parametrize("a_form,w_form", [(AFFINE, AFFINE), (ZP, SYM), (A1, SYM)])
@pytest.mark.parametrize("per_channel", [True, False])
def test_plain_c_oracle_equals_the_torch_restatement(a_form, w_form, per_channel):
    """The second, torch-free statement of the factored product (oracle/fq_oracle.c::orc_code_gemm, what the plain-C
    ABI consumer checks the device against) agrees with oracle/restate.py::code_gemm bit for bit."""
    from oracle import c_oracle
    c_oracle.build()
    qa, qw, ca, cw, m_a, o_a, z_a, m_w, bias = _setup(a_form, w_form, 8, False, per_channel, seed=5)
    for relu in (False, True):
        want = R.code_gemm(ca, cw, m_a, o_a, z_a, m_w, bias, relu=relu)
        got = c_oracle.code_gemm(ca.numpy(), cw.numpy(), float(m_a), float(o_a), float(z_a), m_w.numpy(), bias.numpy(), relu)
        assert torch.equal(torch.from_numpy(got), want)


# ---------------------------------------------------------------------------------------------------------------
# pinned against the reference's own layer outputs (tests/golden/qgemm.npz, minted from the unmodified reference)
# ---------------------------------------------------------------------------------------------------------------
from tests import golden_io  # noqa: E402
from tests.qgemm_golden import Problem  # noqa: E402

QGEMM_GOLDEN = golden_io.load("qgemm")


@pytest.mark.parametrize("name", sorted(QGEMM_GOLDEN))
def test_factored_product_equals_the_reference_layer_output(name):
    """The reference module's eval output (its observers' qparams, its own F.linear / F.conv2d on its own fake-quantised
    tensors) vs oracle/restate.py::code_gemm on the oracle's codes: <= 1e-5 of sum_k |y_a*y_w| (+|bias|)."""
    p = Problem(QGEMM_GOLDEN[name])
    ca, cw, m_a, o_a, z_a, m_w = p.oracle()
    # the codes dequantise to exactly the tensors the reference fed to _forward_func
    ya = (ca - z_a) * m_a + o_a if p.a_form == 2 else ca * m_a + o_a
    assert torch.equal(ya, p.qx) and torch.equal(cw * (m_w.reshape(-1, 1) if m_w.numel() > 1 else m_w), p.qw)
    assert bool((ca == ca.round()).all()) and bool((cw == cw.round()).all())
    out = R.code_gemm(ca, cw, m_a, o_a, z_a, m_w, p.bias)
    err = ((out.double() - p.y.double()).abs() / p.bound().clamp_min(1e-30)).max().item()
    assert err <= 1e-5, err
    assert len(QGEMM_GOLDEN) == 11
