"""Probes of the reference's own behaviour that DESIGN.md relies on.  They import the unmodified reference through
oracle/ref_shim.py and therefore only run where /root/reference exists (the authoring container); skipped elsewhere."""
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present on this machine")


def test_reference_l2norm_pixel_cannot_run():
    """ops.py:237 calls emulate_quantize, which ops.py never imports (ops.py:1-8): quantize_l2norm_pixel raises
    NameError on every input - the reason this one observer type is not provided."""
    ns = ref_shim.load()
    w = torch.randn(8, 4, 3, 3, generator=torch.Generator().manual_seed(5)) * 0.05
    with pytest.raises(NameError, match="emulate_quantize"):
        ns.ops.quantize_l2norm_pixel(w, n_bits=4, signed=True)
    s, o = ns.ops.quantize_minmax_pixel(w, n_bits=4, signed=True)           # its min/max sibling does run
    assert s.shape == (3, 3) and float(o.abs().max()) == 0.0


def test_reference_per_channel_weight_types_crash_in_qbase():
    """modules/base.py:52,128: wt_scale is a [1] Parameter and the observer result is copy_()-ed into it, so a
    per-channel weight observer cannot initialise QBase; ours allocates the scale to fit."""
    ns = ref_shim.load()
    conv = torch.nn.Conv2d(3, 4, 3)
    cls = ns.modules.QConv2d
    m = cls.__new__(cls)
    m.__dict__.update(conv.__dict__)
    m.initialize({"input": {"enable": False, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}},
                  "weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}}})
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 3, 8, 8))
