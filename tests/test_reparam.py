"""BN folding / RepVGG re-parameterisation fused with the per-channel observer.

CPU: the oracle restatement against fixtures minted from the reference's own merge_bn / switch_to_deploy
(tests/golden/reparam.npz, oracle/make_golden.py::golden_reparam), and the host logic.
GPU: `dlmcq_fold_grouped` against the same fixtures - bit-exact - and the module surgery."""
import copy

import pytest
import torch
from torch import nn

from oracle import restate as R
from tests import golden_io as G

CASES = G.load("reparam")
MERGE = sorted(k for k in CASES if k.startswith("merge_bn."))
REPVGG = sorted(k for k in CASES if k.startswith("repvgg."))


def _bn_of(c, tag, eps=None):
    t = tuple(c.inp[f"{tag}_{k}"] for k in ("gamma", "beta", "mean", "var"))
    return t + (eps,) if eps is not None else t


def _oracle_repvgg(c):
    eps = c.meta["eps"]
    bn_id = _bn_of(c, "bnid", eps) if c.meta["has_id"] else None
    return R.repvgg_fuse(c.inp["k3"], _bn_of(c, "bn3", eps), c.inp["k1"], _bn_of(c, "bn1", eps), bn_id, c.meta["groups"])


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", MERGE)
def test_oracle_merge_bn_matches_reference(name):
    c = CASES[name]
    w, b = R.merge_bn_fold(c.inp["w"], c.inp.get("bias"), c.inp["gamma"], c.inp["beta"], c.inp["mean"], c.inp["var"])
    assert G.bits_equal(w, c.out["w"]), G.first_mismatch(w, c.out["w"])
    assert G.bits_equal(b, c.out["bias"]), G.first_mismatch(b, c.out["bias"])
    s, o = R.obs_minmax_channel(w, 8, True, ch_axis=0)
    assert G.bits_equal(s, c.out["obs_scale"]) and G.bits_equal(o.float(), c.out["obs_offset"].float())


@pytest.mark.parametrize("name", REPVGG)
def test_oracle_repvgg_matches_reference(name):
    c = CASES[name]
    w, b = _oracle_repvgg(c)
    assert G.bits_equal(w, c.out["w"]), G.first_mismatch(w, c.out["w"])
    assert G.bits_equal(b, c.out["bias"]), G.first_mismatch(b, c.out["bias"])


def test_fixture_covers_signed_zero_and_tiny_variance():
    c = CASES["repvgg.id"]
    assert (c.inp["bn3_var"] < 1e-6).any() and (c.inp["bn3_gamma"] < 0).any()
    w = c.out["w"]
    assert ((w == 0) & torch.signbit(w)).any() or ((w == 0) & ~torch.signbit(w)).any()


def test_oracle_sqrt_is_ieee_and_torch_cpu_is_within_one_ulp():
    """The reference's std = (var + eps).sqrt() is correctly rounded on CUDA; torch's CPU sqrt (MKL VML) is only
    faithful - see oracle/restate.py::sqrt_ieee.  The oracle pins the IEEE value; the CPU chain stays within 1 ulp."""
    import numpy as np
    x = torch.rand(4096, generator=torch.Generator().manual_seed(1)) * 2 + 0.01
    want = R.sqrt_ieee(x)
    exact = torch.from_numpy(np.sqrt(x.double().numpy()).astype(np.float32))     # innocuous double rounding for sqrt
    assert torch.equal(want, exact)
    ulps = (x.sqrt().view(torch.int32) - want.view(torch.int32)).abs()
    assert int(ulps.max()) <= 1
    probe = torch.tensor([float.fromhex("0x1.17783ep+0")])
    assert R.sqrt_ieee(probe).item() == float.fromhex("0x1.0b7a42p+0")


def test_mapping_functions():
    from dlmc_quant_b200 import reparam as P
    assert P.DEFAULT_CONV_MAPPING_FN("layer1.conv1.1") == "layer1.conv1.0"      # merge_bn.py:14
    assert P.DEFAULT_CONV_MAPPING_FN("layer1.bn1") == "layer1.conv1"            # merge_bn.py:15
    assert P.DEFAULT_CONV_MAPPING_FN("layer1.norm") is None
    assert P.DEFAULT_BN_MAPPING_FN("layer1.conv1.0") == "layer1.conv1.1"
    assert P.DEFAULT_BN_MAPPING_FN("layer1.conv2") == "layer1.bn2"


def test_merge_bn_has_no_cpu_path():
    from dlmc_quant_b200 import _lib
    from dlmc_quant_b200 import reparam as P
    net = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4))
    with pytest.raises(_lib.DlmcqError):
        P.merge_bn(net)


def test_merge_bn_missing_conv_raises_like_reference():
    from dlmc_quant_b200 import reparam as P

    class Odd(nn.Module):
        def __init__(self):
            super().__init__()
            self.norm = nn.BatchNorm2d(4)
    with pytest.raises(ValueError, match="Could not find Conv2d"):
        P.merge_bn(Odd())


# ------------------------------------------------------------------------------------------------ GPU
def _cuda(t):
    return t.cuda().contiguous()


def _merge_entry(c):
    w = _cuda(c.inp["w"])
    return dict(mode="merge_bn", w=w, w_out=torch.empty_like(w), bias_out=torch.empty(w.shape[0], device="cuda"),
                bias=_cuda(c.inp["bias"]) if "bias" in c.inp else None,
                bn=tuple(_cuda(c.inp[k]) for k in ("gamma", "beta", "mean", "var")))


def _repvgg_entry(c):
    eps = c.meta["eps"]
    w = _cuda(c.inp["k3"])
    pack = lambda tag: tuple(_cuda(t) for t in _bn_of(c, tag)) + (eps,)
    return dict(mode="repvgg", w=w, w_out=torch.empty_like(w), bias_out=torch.empty(w.shape[0], device="cuda"),
                w1=_cuda(c.inp["k1"]), bn=pack("bn3"), bn1=pack("bn1"), bn_id=pack("bnid") if c.meta["has_id"] else None)


@pytest.mark.gpu
def test_fold_grouped_is_bit_exact_on_every_fixture_in_one_launch():
    from dlmc_quant_b200 import reparam as P
    names = MERGE + REPVGG
    entries = [_merge_entry(CASES[n]) if n in MERGE else _repvgg_entry(CASES[n]) for n in names]
    stats = P.fold_grouped(entries)
    qp = P.observe_folded(stats, [e["w"].shape[0] for e in entries], 8, True)
    for n, e, (s, o) in zip(names, entries, qp):
        c = CASES[n]
        assert G.bits_equal(e["w_out"].cpu(), c.out["w"]), f"{n}: " + G.first_mismatch(e["w_out"].cpu(), c.out["w"])
        assert G.bits_equal(e["bias_out"].cpu(), c.out["bias"]), f"{n}: " + G.first_mismatch(e["bias_out"].cpu(), c.out["bias"])
        # fused statistics == a second pass of the per-channel observer over the folded weights (ops.py:121-140)
        assert G.bits_equal(s.cpu().reshape(-1), c.out["obs_scale"].reshape(-1)), n
        assert torch.equal(o.cpu().reshape(-1), c.out["obs_offset"].reshape(-1).float()), n


@pytest.mark.gpu
def test_fold_in_place_like_merge_bn():
    from dlmc_quant_b200 import reparam as P
    c = CASES[MERGE[0]]
    e = _merge_entry(c)
    e["w_out"] = e["w"]                               # merge_bn.py:98 writes weight[:] in place
    P.fold_grouped([e])
    assert G.bits_equal(e["w"].cpu(), c.out["w"])


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, padding=1, bias=True)
        self.bn1 = nn.BatchNorm2d(16)
        self.conv2 = nn.Conv2d(16, 24, 3, padding=1, groups=2, bias=False)
        self.bn2 = nn.BatchNorm2d(24)
        self.block = nn.Sequential(nn.Conv2d(24, 8, 1, bias=False), nn.BatchNorm2d(8))

    def forward(self, x):
        return self.block(self.bn2(self.conv2(torch.relu(self.bn1(self.conv1(x))))))


def _randomize(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.weight.copy_(torch.randn(m.num_features, generator=g))
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.5)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.3)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 2 + 0.01)
    return model


@pytest.mark.gpu
def test_merge_bn_model_surgery_and_values():
    from dlmc_quant_b200 import reparam as P
    torch.manual_seed(2333)
    ref = _randomize(_Net(), 5).eval()
    net = copy.deepcopy(ref).cuda()
    x = torch.randn(2, 3, 12, 12)
    want = ref(x)
    merged, stats = P.merge_bn(net, return_stats=True)
    assert merged is net                              # inplace=False modifies the argument (merge_bn.py:61-62)
    assert isinstance(net.bn1, nn.Identity) and isinstance(net.bn2, nn.Identity) and isinstance(net.block[1], nn.Identity)
    assert net.conv2.bias is not None
    for cn, bnn in (("conv1", "bn1"), ("conv2", "bn2"), ("block.0", "block.1")):
        from operator import attrgetter
        conv, bn = attrgetter(cn)(ref), attrgetter(bnn)(ref)
        w, b = R.merge_bn_fold(conv.weight.data, None if conv.bias is None else conv.bias.data, bn.weight.data,
                               bn.bias.data, bn.running_mean.data, bn.running_var.data)
        got = attrgetter(cn)(net)
        assert G.bits_equal(got.weight.data.cpu(), w) and G.bits_equal(got.bias.data.cpu(), b), cn
        s, _ = R.obs_minmax_channel(w, 4, True, ch_axis=0)
        (s_got, _), = P.observe_folded(stats[cn], [w.shape[0]], 4, True)
        assert G.bits_equal(s_got.cpu(), s), cn
    got = net(x.cuda()).cpu()
    # BN eps (1e-5) vs the fold's 1e-7: equal only to a tolerance, as in the reference
    assert torch.allclose(got, want, rtol=2e-3, atol=2e-3)


class _Branch(nn.Sequential):
    def __init__(self, cin, cout, k, stride, pad, groups):
        super().__init__()
        self.add_module("conv", nn.Conv2d(cin, cout, k, stride, pad, groups=groups, bias=False))
        self.add_module("bn", nn.BatchNorm2d(cout))


class _Block(nn.Module):
    """Train-form RepVGG block with the attribute names repvgg_model_convert looks for."""

    def __init__(self, cin, cout, stride=1, groups=1):
        super().__init__()
        self.rbr_identity = nn.BatchNorm2d(cin) if cin == cout and stride == 1 else None
        self.rbr_dense = _Branch(cin, cout, 3, stride, 1, groups)
        self.rbr_1x1 = _Branch(cin, cout, 1, stride, 0, groups)

    def forward(self, x):
        if hasattr(self, "rbr_reparam"):
            return torch.relu(self.rbr_reparam(x))
        ident = 0 if self.rbr_identity is None else self.rbr_identity(x)
        return torch.relu(self.rbr_dense(x) + self.rbr_1x1(x) + ident)


@pytest.mark.gpu
def test_repvgg_model_convert_matches_oracle_and_keeps_the_function():
    from dlmc_quant_b200 import reparam as P
    torch.manual_seed(2333)
    ref = _randomize(nn.Sequential(_Block(3, 16, 2), _Block(16, 16), _Block(16, 16, groups=4), _Block(16, 32, 2)), 9).eval()
    x = torch.randn(2, 3, 32, 32)
    want = ref(x)
    deploy, stats = P.repvgg_model_convert(copy.deepcopy(ref).cuda(), return_stats=True)
    for i, blk in enumerate(ref):
        pack = lambda bn: (bn.weight.data, bn.bias.data, bn.running_mean.data, bn.running_var.data, bn.eps)
        w, b = R.repvgg_fuse(blk.rbr_dense.conv.weight.data, pack(blk.rbr_dense.bn), blk.rbr_1x1.conv.weight.data,
                             pack(blk.rbr_1x1.bn), pack(blk.rbr_identity) if blk.rbr_identity is not None else None,
                             blk.rbr_dense.conv.groups)
        d = deploy[i]
        assert d.deploy and not hasattr(d, "rbr_dense") and not hasattr(d, "rbr_1x1") and not hasattr(d, "rbr_identity")
        assert G.bits_equal(d.rbr_reparam.weight.data.cpu(), w), f"block {i}"
        assert G.bits_equal(d.rbr_reparam.bias.data.cpu(), b), f"block {i}"
        s, _ = R.obs_minmax_channel(w, 8, True, ch_axis=0)
        (s_got, _), = P.observe_folded(stats[str(i)], [w.shape[0]], 8, True)
        assert G.bits_equal(s_got.cpu(), s)
    assert torch.allclose(deploy(x.cuda()).cpu(), want, rtol=1e-3, atol=1e-3)


# ------------------------------------------------------------------------------------------------ plain-C oracle
@pytest.mark.parametrize("name", MERGE + REPVGG)
def test_c_oracle_fold_matches_reference(name):
    """The torch-free C restatement (IEEE sqrtf) against the fixtures minted from the reference."""
    from oracle import c_oracle
    c = CASES[name]
    n = lambda t: t.numpy()
    if name in MERGE:
        w, b = c_oracle.merge_bn(n(c.inp["w"]), n(c.inp["bias"]) if "bias" in c.inp else None, n(c.inp["gamma"]),
                                 n(c.inp["beta"]), n(c.inp["mean"]), n(c.inp["var"]))
    else:
        pack = lambda tag: tuple(n(t) for t in _bn_of(c, tag))
        w, b = c_oracle.repvgg_fuse(n(c.inp["k3"]), pack("bn3"), n(c.inp["k1"]), pack("bn1"),
                                    pack("bnid") if c.meta["has_id"] else None, c.meta["eps"])
    assert G.bits_equal(torch.from_numpy(w), c.out["w"]), G.first_mismatch(torch.from_numpy(w), c.out["w"])
    assert G.bits_equal(torch.from_numpy(b), c.out["bias"])
