"""GPU parity tests of the fused BatchNorm (+ residual) (+ ReLU) -> input fake-quant kernels (SURVEY.md 8f row f2;
csrc/bnq_kernels.cu through the C ABI dlmcq_bnq_forward / dlmcq_bnq_backward) and of the model rewiring in
dlmc_quant_b200/fuse.py.

What is compared with what:
  * BatchNorm stage  vs torch.nn.functional.batch_norm (the library op the reference's models call): floating-point
    reduction work, tolerance stated below (1e-5 relative to the tensor's scale) - two batch-norm implementations
    never agree bit for bit;
  * quantizer stage  vs this repo's stand-alone kernel on the fused op's own plain output: BIT-EXACT (and that
    kernel is bit-exact against the reference chain, tests/test_gpu_fq.py);
  * backward         stage by stage on identical inputs: fake-quant backward of `a` (oracle chain + autograd),
    ReLU mask, then torch's batch-norm backward driven with the same dz - no tie can flip between the two sides;
  * rewiring         with the fused kernels disabled the rewired model is bit-identical to the original one."""
import copy

import pytest
import torch
import torch.nn.functional as TF

from oracle import restate as R

pytestmark = pytest.mark.gpu

AFFINE = 1


def F():
    from dlmc_quant_b200 import functional
    return functional


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t.contiguous()


class _BN:
    """Minimal stand-in for nn.BatchNorm2d state."""

    def __init__(self, c, training, affine=True, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.weight = (torch.rand(c, generator=g) + 0.5).cuda() if affine else None
        self.bias = (torch.randn(c, generator=g) * 0.3).cuda() if affine else None
        self.running_mean = (torch.randn(c, generator=g) * 0.1).cuda()
        self.running_var = (torch.rand(c, generator=g) + 0.5).cuda()
        self.training, self.momentum, self.eps, self.track_running_stats = training, 0.1, 1e-5, True
        self.num_batches_tracked = torch.zeros((), dtype=torch.long).cuda()

    def clone(self):
        b = copy.copy(self)
        for n in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            v = getattr(self, n)
            setattr(b, n, v.clone() if v is not None else None)
        return b

    def __call__(self, x):
        return TF.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, self.training,
                             self.momentum, self.eps)


SHAPES = [(8, 64, 14, 14), (4, 256, 7, 7), (3, 2048, 5, 5), (2, 48, 9, 9), (16, 8, 3, 3), (5, 1028, 2, 3), (37, 128)]


def _inputs(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = _cl((torch.randn(shape, generator=g) * 1.7 + 0.4).cuda().to(dtype))
    idn = _cl(torch.relu(torch.randn(shape, generator=g)).cuda().to(dtype))
    return x, idn


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("relu,resid,quant", [(True, False, True), (True, True, True), (False, False, False),
                                              (True, False, False), (False, True, False), (False, False, True)])
def test_fused_forward_stages(shape, training, relu, resid, quant):
    from dlmc_quant_b200 import fuse
    x, idn = _inputs(shape, 2333)
    bn = _BN(shape[1], training)
    ref_bn = bn.clone()
    scale = torch.tensor([0.21], device="cuda")
    offset = torch.tensor([0.0 if relu else -1.3], device="cuda")
    lo, hi = 0, 15
    g = R.lsq_g(x.numel(), hi)
    state = (bn.running_mean, bn.running_var, training, bn.momentum, bn.eps)
    a, aq = fuse.BnActQuantFunction.apply(x, bn.weight, bn.bias, scale if quant else None, idn if resid else None, state,
                                          (offset, lo, hi, g) if quant else None, relu, True)
    # stage 1: BatchNorm (+identity) (+ReLU) against the library op
    a_ref, _ = R.bn_act_fq_chain(x, ref_bn.weight, ref_bn.bias, ref_bn.running_mean, ref_bn.running_var, training,
                                 ref_bn.momentum, ref_bn.eps, idn if resid else None, relu, None, None, lo, hi, g)
    tol = 1e-5 * float(a_ref.abs().max()) + 1e-6
    assert float((a - a_ref).abs().max()) <= tol, (float((a - a_ref).abs().max()), tol)
    assert a.stride() == x.stride()
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-5, atol=1e-6)
    # stage 2: the quantizer is the stand-alone kernel's arithmetic, bit for bit
    if quant:
        want = F().fq_forward(a, scale, offset, lo, hi, AFFINE, g=g)
        assert torch.equal(aq.view(torch.int32), want.view(torch.int32)), "fused a_q != fq_forward(fused a)"
        # and the whole chain against the oracle chain, away from rounding ties the BatchNorm tolerance can flip
        _, aq_ref = R.bn_act_fq_chain(x, ref_bn.weight, ref_bn.bias, None, None, True, 0.0, ref_bn.eps,
                                      idn if resid else None, relu, scale, offset, lo, hi, g) if training else \
            R.bn_act_fq_chain(x, ref_bn.weight, ref_bn.bias, bn.running_mean, bn.running_var, False, 0.0, ref_bn.eps,
                              idn if resid else None, relu, scale, offset, lo, hi, g)
        flips = (aq != aq_ref)
        assert float(flips.float().mean()) < 2e-4, float(flips.float().mean())
        assert float((aq - aq_ref).abs().max()) <= float(scale) * 1.0001 + tol     # a flip is exactly one step


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("relu,resid,quant,plain_grad", [(True, False, True, False), (True, True, True, True),
                                                         (True, False, False, True), (False, True, False, True),
                                                         (True, False, True, True), (False, False, True, False)])
def test_fused_backward_stages(shape, training, relu, resid, quant, plain_grad):
    from dlmc_quant_b200 import fuse
    x, idn = _inputs(shape, 77)
    x.requires_grad_(True)
    idn.requires_grad_(resid)
    bn = _BN(shape[1], training, seed=3)
    bn.weight.requires_grad_(True)
    bn.bias.requires_grad_(True)
    scale = torch.tensor([0.19], device="cuda", requires_grad=True)
    offset = torch.tensor([0.0 if relu else -1.1], device="cuda")
    lo, hi = 0, 15
    g = R.lsq_g(x.numel(), hi)
    state = (bn.running_mean.clone(), bn.running_var.clone(), training, bn.momentum, bn.eps)
    a, aq = fuse.BnActQuantFunction.apply(x, bn.weight, bn.bias, scale if quant else None, idn if resid else None, state,
                                          (offset, lo, hi, g) if quant else None, relu, True)
    gen = torch.Generator().manual_seed(5)
    d_a = _cl(torch.randn(shape, generator=gen).cuda()) if plain_grad else None
    d_q = _cl(torch.randn(shape, generator=gen).cuda()) if quant else None
    outs, grads = [], []
    if plain_grad:
        outs.append(a); grads.append(d_a)
    if quant:
        outs.append(aq); grads.append(d_q)
    ins = [x, bn.weight, bn.bias] + ([scale] if quant else []) + ([idn] if resid else [])
    got = torch.autograd.grad(outs, ins, grads)
    dx, dgamma, dbeta = got[0], got[1], got[2]
    ds = got[3] if quant else None
    did = got[-1] if resid else None

    # reference, stage by stage on the SAME plain output a
    a0 = a.detach()
    da = torch.zeros_like(a0)
    ds_ref = None
    if quant:
        a1 = a0.clone().requires_grad_(True)
        s1 = scale.detach().clone().requires_grad_(True)
        yq = R.fq_affine(a1, s1, offset, lo, hi, g)
        assert torch.equal(yq.detach(), aq.detach())
        da_q, ds_ref = torch.autograd.grad(yq, [a1, s1], d_q)
        da = da + da_q
    if plain_grad:
        da = da + d_a
    dz = da * (a0 > 0) if relu else da
    x2 = x.detach().clone().requires_grad_(True)
    w2 = bn.weight.detach().clone().requires_grad_(True)
    b2 = bn.bias.detach().clone().requires_grad_(True)
    z = TF.batch_norm(x2, state[0].clone(), state[1].clone(), w2, b2, training, bn.momentum, bn.eps)
    dx_ref, dg_ref, db_ref = torch.autograd.grad(z, [x2, w2, b2], dz)

    def close(mine, ref, what, rel=2e-5):
        tol = rel * float(ref.abs().max()) + 1e-7
        err = float((mine - ref).abs().max())
        assert err <= tol, (what, err, tol)
    close(dx, dx_ref, "dx", rel=5e-5)              # dx subtracts the two channel means: a few ulp of the largest term
    n = x.numel() / shape[1]
    close(dgamma, dg_ref, "dgamma", rel=1e-5 * max(1.0, n ** 0.5 / 8))
    close(dbeta, db_ref, "dbeta", rel=1e-5 * max(1.0, n ** 0.5 / 8))
    if resid:
        # d identity IS dz; the oracle chain's own da is (d_q * s) / s - a 1-ulp wobble
        # (where da_q and d_a nearly cancel, that ulp is an absolute, not a relative, error: atol from |d_q|)
        assert torch.allclose(did, dz, rtol=1e-6, atol=2e-7 * float(d_q.abs().max()) if quant else 0), "d identity must be dz"
    if quant:
        floor = 4e-7 * float(d_q.abs().sum()) * hi * g
        assert abs(float(ds) - float(ds_ref)) <= 1e-5 * abs(float(ds_ref)) + floor, (float(ds), float(ds_ref))


def test_fused_bf16_and_determinism():
    from dlmc_quant_b200 import fuse
    x, idn = _inputs((8, 64, 12, 12), 9, torch.bfloat16)
    bn = _BN(64, True)
    scale, offset = torch.tensor([0.25], device="cuda"), torch.zeros(1, device="cuda")
    g = R.lsq_g(x.numel(), 15)
    outs = []
    for _ in range(2):
        state = (bn.running_mean.clone(), bn.running_var.clone(), True, 0.1, 1e-5)
        xx = x.clone().requires_grad_(True)
        a, aq = fuse.BnActQuantFunction.apply(xx, bn.weight, bn.bias, scale, idn, state, (offset, 0, 15, g), True, True)
        (dx,) = torch.autograd.grad([a, aq], [xx], [torch.ones_like(a), torch.ones_like(aq)])
        outs.append((a, aq, dx))
    for u, v in zip(*outs):
        assert torch.equal(u, v)                                   # run-to-run deterministic
    a, aq, _ = outs[0]
    ref = torch.relu(TF.batch_norm(x.float(), None, None, bn.weight, bn.bias, True, 0.1, 1e-5) + idn.float())
    assert float((a.detach().float() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())
    want = F().fq_forward(a, scale, offset, 0, 15, AFFINE, g=g)
    assert torch.equal(aq.view(torch.int16), want.view(torch.int16))


def test_fallbacks_and_errors():
    from dlmc_quant_b200 import fuse
    bn = torch.nn.BatchNorm2d(6).cuda()
    x = torch.randn(2, 6, 4, 4, device="cuda")                     # NCHW-contiguous and C % 4 != 0: unfused path
    assert not fuse.fusable(x, bn)
    a, aq = fuse.bn_act_quant(x, bn, None, None, True, True)
    assert aq is None and torch.equal(a, torch.relu(TF.batch_norm(x, None, None, bn.weight, bn.bias, True)))
    bn8 = torch.nn.BatchNorm2d(8, momentum=None).cuda()
    assert not fuse.fusable(_cl(torch.randn(2, 8, 4, 4, device="cuda")), bn8)


CFG = {"weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}},
       "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}},
       "exclude_layers": [], "override_options": [], "momentum": 0.1}


def _resnet(kind):
    import torchvision
    torch.manual_seed(2333)
    if kind == "tv50":
        return torchvision.models.resnet50(num_classes=10)
    if kind == "tv18":
        return torchvision.models.resnet18(num_classes=10)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["tv18", "tv50"])
def test_rewired_model_is_identical_when_kernels_are_disabled_and_close_when_enabled(kind, monkeypatch):
    """(1) Wiring: with the fused kernels switched off the rewired model IS the original model, bit for bit.
    (2) Kernels on: every residual block, fed the reference model's own input to that block (a deep random-initialised
    W4A4 network amplifies a single flipped code chaotically - profiles/debug_fuse.py - so blocks are compared one at
    a time), reproduces the reference block's output up to BatchNorm rounding: almost all elements within 1e-4 of the
    tensor's scale, the rest (code flips downstream of a tie) bounded; same for two consecutive blocks through the
    pre-quantised hand-off, and for the block's gradients."""
    from dlmc_quant_b200 import fuse, quantize_model
    base = _resnet(kind).cuda().to(memory_format=torch.channels_last)
    quantize_model(base, copy.deepcopy(CFG), None)
    x = _cl(torch.randn(8, 3, 64, 64, device="cuda"))
    t = torch.randint(0, 10, (8,), device="cuda")
    base.train()
    with torch.no_grad():
        base(x)                                                    # lazy observer init
    fused = copy.deepcopy(base)
    h = fuse.fuse_bn_act_quant(fused)
    assert h.blocks == (8 if kind == "tv18" else 16) and h.batchnorms >= 1

    def step(m):
        m.zero_grad(set_to_none=True)
        y = m(x)
        TF.cross_entropy(y, t).backward()
        return y.detach(), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

    monkeypatch.setattr(fuse, "fusable", lambda *a, **k: False)
    off = copy.deepcopy(base)
    h_off = fuse.fuse_bn_act_quant(off)
    ref = copy.deepcopy(base)
    y0, g0 = step(ref)
    y1, g1 = step(off)
    assert torch.equal(y0, y1) and g0.keys() == g1.keys()
    bad = [n for n in g0 if not torch.allclose(g0[n], g1[n], rtol=1e-3, atol=1e-6)]    # cuDNN wgrad is not bit-reproducible
    assert not bad, bad
    for (n0, b0), (n1, b1) in zip(ref.named_buffers(), off.named_buffers()):
        assert n0 == n1 and torch.equal(b0, b1), n0
    h_off.unfuse()
    monkeypatch.undo()

    # BatchNorm buffers after ONE training forward of the whole model from the same state: batch counters agree
    # everywhere, running statistics agree in the stem and the first block (deeper ones see chaotically different inputs)
    ref3, fused3 = copy.deepcopy(base), copy.deepcopy(base)
    fuse.fuse_bn_act_quant(fused3)
    with torch.no_grad():
        ref3(x), fused3(x)
    for (n0, b0), (n1, b1) in zip(ref3.named_buffers(), fused3.named_buffers()):
        if "num_batches" in n0:
            assert torch.equal(b0, b1), n0
        elif "running" in n0 and (n0.startswith("bn1") or n0.startswith("layer1.0")):
            assert torch.allclose(b0, b1, rtol=1e-3, atol=1e-4), n0

    # (2) block by block on the reference's own block inputs
    ref2 = copy.deepcopy(base)
    io = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, n=n: io.__setitem__(n, (i[0].detach(), o.detach())))
             for n, m in ref2.named_modules() if n.startswith("layer") and n.count(".") == 1]
    with torch.no_grad():
        ref2(x)
    for hk in hooks:
        hk.remove()
    names = list(io)
    fblocks = dict(fused.named_modules())
    rblocks = dict(ref2.named_modules())

    def check(got, want, what, frac_tol, max_steps=3.0):
        scale = float(want.abs().max())
        d = (got - want).abs()
        frac = float((d > 1e-4 * scale).float().mean())
        assert frac <= frac_tol, (what, frac)
        assert float(d.max()) <= 0.5 * scale, (what, float(d.max()), scale)

    with torch.no_grad():
        for n in names:
            inp, out = io[n]
            check(fblocks[n](inp), out, n, 0.03)
        for n0, n1 in zip(names, names[1:]):
            if n0.split(".")[0] != n1.split(".")[0]:
                continue                                          # hand-off within a stage
            mid = fblocks[n0](io[n0][0])
            assert getattr(mid, "_dlmcq_q", None) is not None and mid._dlmcq_q[0] is fblocks[n1].conv1
            check(fblocks[n1](mid), io[n1][1], n0 + "->" + n1, 0.15)
    # gradients of one block, same input and upstream gradient on both sides
    n = names[1]
    inp, out = io[n]
    gy = torch.randn_like(out)
    grads = []
    for blk in (rblocks[n], fblocks[n]):
        blk.zero_grad(set_to_none=True)
        xi = inp.clone().requires_grad_(True)
        blk(xi).backward(gy)
        grads.append([xi.grad] + [p.grad for p in blk.parameters()])
    for a_, b_ in zip(*grads):
        a_, b_ = a_.flatten().double(), b_.flatten().double()
        if float(a_.norm()) > 0:
            cos = float(torch.dot(a_, b_) / (a_.norm() * b_.norm()))
            assert cos > 0.995, cos
    # eval mode uses the running buffers: first block, same input, same buffers (the block-level calls above
    # updated the fused model's running statistics more often than the reference's)
    fused.load_state_dict(ref2.state_dict())
    fused.eval(); ref2.eval()
    with torch.no_grad():
        inp, _ = io[names[0]]
        check(fblocks[names[0]](inp), rblocks[names[0]](inp), "eval " + names[0], 0.03)
    h.unfuse()
    with torch.no_grad():
        assert torch.equal(fused(x), fused(x)) and 'forward' not in fused.layer1[0].__dict__


def test_fused_sites_actually_run_fused(monkeypatch):
    """Count the fused launches of one ResNet-50 forward: 49 of the 53 BatchNorms carry a quantizer (conv2/conv3 inputs
    and the next block's conv1), the rest run BatchNorm(+ReLU) alone; nothing falls back."""
    from dlmc_quant_b200 import fuse, quantize_model
    m = _resnet("tv50").cuda().to(memory_format=torch.channels_last)
    quantize_model(m, copy.deepcopy(CFG), None)
    x = _cl(torch.randn(4, 3, 64, 64, device="cuda"))
    m.train()
    with torch.no_grad():
        m(x)
    fuse.fuse_bn_act_quant(m)
    calls = {"quant": 0, "plain": 0, "unfused": 0}
    orig = fuse.BnActQuantFunction.apply

    def spy(*args):
        calls["quant" if args[6] is not None else "plain"] += 1
        return orig(*args)
    monkeypatch.setattr(fuse.BnActQuantFunction, "apply", staticmethod(spy))
    monkeypatch.setattr(fuse, "_unfused", lambda *a: (_ for _ in ()).throw(AssertionError("fell back to the unfused chain")))
    m(x)
    assert calls["quant"] == 16 * 2 + 15 and calls["plain"] == 53 - calls["quant"], calls
