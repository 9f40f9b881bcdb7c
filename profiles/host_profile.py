"""Host-side (Python) cost of one fused QAT step: cProfile over 10 steps of ResNet-50 channels_last batch 128 -
which functions the ~22 ms of host time per step go to."""
import copy
import cProfile
import os
import pstats
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torchvision  # noqa: E402

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
arm = sys.argv[1] if len(sys.argv) > 1 else "ours_fused"
model = torchvision.models.resnet50().to(dev).to(memory_format=torch.channels_last)
if arm != "fp32":
    from dlmc_quant_b200 import quantize_model
    quantize_model(model, copy.deepcopy(bench.QAT_CFG), None)
x = torch.randn(128, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 1000, (128,), device=dev)
model.train()
with torch.no_grad():
    model(x[:8])
if arm != "fp32":
    from dlmc_quant_b200.fuse import fuse_bn_act_quant
    from dlmc_quant_b200.quantize import group_weight_quantizers
    group_weight_quantizers(model)
    if arm == "ours_fused":
        fuse_bn_act_quant(model)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=5e-4)
crit = nn.CrossEntropyLoss()


def step():
    opt.zero_grad()
    loss = crit(model(x), t)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
