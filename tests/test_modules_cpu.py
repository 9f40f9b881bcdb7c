"""CPU tests of the host-side mirror of the reference interface: class swap, state names and shapes
(checkpoint / fnmatch compatibility), config handling, and that compute fails loudly without CUDA."""
import copy
import os
from fnmatch import fnmatch

import pytest
import torch
from torch import nn

from dlmc_quant_b200 import quantize_model
from dlmc_quant_b200.quantize import get_layers
from dlmc_quant_b200.scalar import FSPTQuant, RootQ, modules
from dlmc_quant_b200.scalar.utils import get_qrange, infer_ch_axis
from tests.golden_io import load

CFG = {"weight": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": True}},
       "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}},
       "exclude_layers": [], "override_options": [], "momentum": 0.1}


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 8, 3, padding=1)
        self.block = nn.Sequential(nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.ReLU(), nn.Conv2d(8, 4, 1))
        self.fc = nn.Linear(4, 10)

    def forward(self, x):
        return self.fc(self.block(self.conv1(x)).mean((2, 3)))


def test_get_qrange_matches_reference():
    assert get_qrange(True, 4) == (-7, 7) and get_qrange(False, 4) == (0, 15)
    assert get_qrange(True, 8) == (-127, 127) and get_qrange(False, 8) == (0, 255)


def test_quantize_model_swaps_in_place_and_keeps_weights():
    net = Net()
    w = net.block[0].weight
    quantize_model(net, copy.deepcopy(CFG), None)
    assert type(net.conv1) is modules.QConv2d and type(net.fc) is modules.QLinear
    assert type(net.block[0]) is modules.QConv2d and net.block[0].weight is w
    assert net.block[0].in_max_val == 15 and net.block[0].wt_min_val == -7


def test_exclude_and_override_options():
    cfg = copy.deepcopy(CFG)
    cfg["exclude_layers"] = ["conv1", "fc"]
    cfg["override_options"] = [{"layers": ["block.2"], "options": {"weight": {"args": {"n_bits": 8}}}}]
    net = Net()
    quantize_model(net, cfg, None)
    assert type(net.conv1) is nn.Conv2d and type(net.fc) is nn.Linear
    assert net.block[2].wt_max_val == 127 and net.block[0].wt_max_val == 7
    assert get_layers(net, filter_types=(nn.Conv2d,)) == ["conv1", "block.0", "block.2"]


QBASE, ROOTQ, FSPTQ = load("qbase"), load("rootq"), load("fsptq")


def _state_names(case, prefix):
    return sorted(k[len(prefix):] for k in case.out if k.startswith(prefix))


def test_qbase_state_names_match_reference():
    net = Net()
    quantize_model(net, copy.deepcopy(CFG), None)
    m = net.block[0]
    params = {n: tuple(p.shape) for n, p in m.named_parameters()}
    assert params["in_scale"] == (1,) and params["wt_scale"] == (1,)
    ref = QBASE["conv_mm_w4a4"]
    assert set(_state_names(ref, "param_")) <= set(params)
    bufs = dict(m.named_buffers())
    assert set(bufs) == {"in_init_state", "wt_init_state"}          # offsets are None until first forward
    assert m.in_offset is None and m.wt_offset is None
    assert {"in_offset", "wt_offset"} | set(bufs) == set(_state_names(ref, "buf_"))


def test_rootq_state_names_shapes_and_filters():
    net = Net()
    quantize_model(net, copy.deepcopy(CFG), None, quantization_type="RootQ")
    m = net.block[0]
    assert type(m) is RootQ.RootQConv2d and m.momentum == 0.1
    ref = ROOTQ["conv_w4a4_step1"]
    for n in _state_names(ref, "param_"):
        assert tuple(getattr(m, n).shape) == tuple(ref.out["param_" + n].shape) == ()
    for n in _state_names(ref, "buf_"):
        assert tuple(getattr(m, n).shape) == ()
    assert float(m.wt_alpha) == 0.25 and float(m.wt_upper) == 3.0 and float(m.wt_lower) == -4.0
    names = [n for n, _ in net.named_parameters()]
    # trainer/quantization_aware_training_trainer.py:92,139 and RootQ_train.py:69 filter on these patterns
    assert any(fnmatch(n, "*wt_alpha*") for n in names) and any(fnmatch(n, "*in_scale*") for n in names)
    assert any(n.endswith("upper") for n in names) and any(n.endswith("lower") for n in names)


def test_fsptq_state_and_api():
    cfg = copy.deepcopy(CFG)
    cfg["weight"] = {"enable": True, "type": "minmax_channel", "recon_type": "adaround",
                     "args": {"n_bits": 4, "signed": True, "ch_axis": 0}}
    net = Net()
    quantize_model(net, cfg, None, quantization_type="FSPTQ")
    m, lin = net.block[0], net.fc
    assert type(m) is FSPTQuant.FSPTQConv2d and type(lin) is FSPTQuant.FSPTQLinear
    assert tuple(m.wt_scale.shape) == (8, 1, 1, 1) and tuple(lin.wt_scale.shape) == (10, 1)
    assert tuple(m.in_scale.shape) == (1,) and tuple(m.in_offset.shape) == (1,)
    assert tuple(m.alpha.shape) == tuple(m.weight.shape) and (m.gamma, m.zeta) == (-0.1, 1.1)
    assert torch.equal(m.org_weight, m.weight.detach())
    ref = FSPTQ["conv_ada_w4a8"]
    assert set(_state_names(ref, "param_")) <= {n for n, _ in m.named_parameters()}
    m.change_quant_state(False, True)
    assert (m.wt_quant, m.act_quant) == (False, True)
    m.reinit_parameters()
    assert any(fnmatch(n, "*scale*") for n, _ in net.named_parameters())      # FSPTQuant.py:90


def test_state_dict_round_trip_resets_host_flags():
    net = Net()
    quantize_model(net, copy.deepcopy(CFG), None)
    m = net.block[0]
    sd = net.state_dict()
    sd["block.0.in_init_state"] = torch.ones(1)
    m._host_init = {"in": False, "wt": False}
    net.load_state_dict(sd)
    assert m._host_init == {"in": None, "wt": None} and float(m.in_init_state) == 1.0


def test_per_channel_scales_are_sized_at_initialize_and_load_from_checkpoints():
    """Per-channel types: the scale Parameters get their final shape in initialize() (the reference registers [1] and
    crashes in copy_, base.py:116,128), so an optimizer built before the first forward holds the right objects; a
    calibrated checkpoint (per-channel scales, offset buffers present) loads strictly into a fresh model."""
    cfg = copy.deepcopy(CFG)
    cfg["weight"] = {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}}
    cfg["input"] = {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": False}}
    net = Net()
    quantize_model(net, copy.deepcopy(cfg), None)
    assert tuple(net.conv1.wt_scale.shape) == (8, 1, 1, 1) and tuple(net.conv1.in_scale.shape) == (1, 3, 1, 1)
    assert tuple(net.fc.wt_scale.shape) == (10, 1) and tuple(net.fc.in_scale.shape) == (1, 4)
    # a calibrated state dict from a model whose placeholders were [1]
    net2 = Net()
    quantize_model(net2, copy.deepcopy(CFG), None)
    sd = net.state_dict()
    for name, mod in net.named_modules():
        if isinstance(mod, modules.QBase):
            sd[name + ".in_offset"] = torch.zeros_like(mod.in_scale)
            sd[name + ".wt_offset"] = torch.zeros_like(mod.wt_scale)
    net2.load_state_dict(sd, strict=True)
    assert tuple(net2.conv1.wt_scale.shape) == (8, 1, 1, 1) and tuple(net2.conv1.wt_offset.shape) == (8, 1, 1, 1)
    assert tuple(net2.fc.in_offset.shape) == (1, 4)


def test_channel_axis_inference():
    w = torch.zeros(8, 4, 3, 3)
    assert infer_ch_axis(w, torch.ones(1)) is None and infer_ch_axis(w, torch.tensor(1.0)) is None
    assert infer_ch_axis(w, torch.ones(8, 1, 1, 1)) == (0, 0)
    assert infer_ch_axis(w, torch.ones(1, 4, 1, 1)) == (1, 1)
    assert infer_ch_axis(w, torch.ones(3, 3)) == (2, 3)
    assert infer_ch_axis(torch.zeros(5, 7), torch.ones(5, 1)) == (0, 0)


def test_forward_without_cuda_raises():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dlmc_quant_b200._lib import DlmcqError
    net = Net()
    quantize_model(net, copy.deepcopy(CFG), None)
    with pytest.raises(DlmcqError):
        net(torch.randn(2, 3, 8, 8))


def test_compat_aliases_resolve_reference_import_paths():
    """`import dlmc.quantization.scalar...` (the reference's paths, as used by its trainers) -> this package."""
    import subprocess
    import sys
    code = (
        "import dlmc_quant_b200.compat as c; c.install();"
        "from dlmc.utils.quantize import quantize_model;"
        "from dlmc.quantization.scalar import modules as qnn, RootQ as RQ, FSPTQuant as FSPQ;"
        "from dlmc.quantization.scalar.ops import get_qparams_tensor;"
        "from dlmc.quantization.scalar.utils import get_qrange;"
        "from dlmc.utils.merge_bn import merge_bn;"
        "import dlmc_quant_b200.scalar.modules as ours;"
        "assert qnn.QConv2d is ours.QConv2d and RQ.RootQConv2d.__module__.startswith('dlmc_quant_b200');"
        "assert get_qrange(True, 4) == (-7, 7) and quantize_model.__module__ == 'dlmc_quant_b200.quantize';"
        "print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_weight_group_partition_is_contiguous_and_balanced():
    """group_weight_quantizers(model, n_groups): forward-ordered, contiguous chunks of about equal parameter bytes
    (so that data-parallel gradient buckets fill while the backward pass is still running)."""
    from dlmc_quant_b200 import quantize_model
    from dlmc_quant_b200.quantize import group_weight_quantizers
    net = torch.nn.Sequential(*[torch.nn.Conv2d(8, 8, 3) for _ in range(6)], torch.nn.Conv2d(8, 64, 3))
    quantize_model(net, copy.deepcopy(CFG), None)
    handle = group_weight_quantizers(net, n_groups=3)
    chunks = [g._candidates for g in handle.groups]
    assert [m for c in chunks for m in c] == list(net) and all(chunks)      # forward order, nothing lost
    sizes = [sum(m.weight.numel() for m in c) for c in chunks]
    assert 2 <= len(chunks) <= 3 and max(sizes) <= 0.7 * sum(sizes)         # bytes spread over the groups
    assert len(group_weight_quantizers(net).groups) == 1        # single process: one group
    handle.remove()
