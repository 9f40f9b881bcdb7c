"""Where do a fused and an unfused ResNet-50 diverge?  Per-block relative difference of the training-mode forward
(hooks on every residual block): BatchNorm rounding flips a few codes per layer; this shows how that grows with depth."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlmc_quant_b200 import fuse, quantize_model  # noqa: E402

CFG = {"weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}},
       "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}},
       "exclude_layers": [], "override_options": [], "momentum": 0.1}
import torchvision
torch.manual_seed(2333)
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
CFG["weight"]["args"]["n_bits"] = CFG["input"]["args"]["n_bits"] = bits
m = torchvision.models.resnet50(num_classes=10).cuda().to(memory_format=torch.channels_last)
quantize_model(m, copy.deepcopy(CFG), None)
x = torch.randn(8, 3, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last)
m.train()
with torch.no_grad():
    m(x)
f = copy.deepcopy(m)
fuse.fuse_bn_act_quant(f)
outs = {}


def hook(tag):
    def h(mod, inp, out):
        outs.setdefault(tag, []).append(out.detach().clone())
    return h


for (n0, b0), (n1, b1) in zip(m.named_modules(), f.named_modules()):
    if n0.count(".") == 1 and n0.startswith("layer"):
        b0.register_forward_hook(hook(n0))
        b1.register_forward_hook(hook(n0))
with torch.no_grad():
    y0, y1 = m(x), f(x)
for k, (a, b) in outs.items():
    d = (a - b).abs()
    print(f"{k:12s} max|a|={float(a.abs().max()):9.4f} max diff={float(d.max()):9.5f} mean diff={float(d.mean()):.3e} "
          f"frac differing={float((d > 1e-4 * a.abs().max()).float().mean()):.4f}")
print("logits", float((y0 - y1).abs().max()), float(y0.abs().max()))
