"""Where the QAT step's GPU time goes: torch.profiler kernel table of 3 steps of the `ours` arm of
profiles/qat_images_per_s.py (ResNet-50 W4A4, batch 128).   python profiles/qat_kernel_breakdown.py [arm]"""
import copy
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import qat_images_per_s as Q  # noqa: E402

arm = sys.argv[1] if len(sys.argv) > 1 else "ours"
import torchvision
torch.backends.cudnn.benchmark = True
torch.manual_seed(2333)
device = torch.device("cuda", 0)
model = torchvision.models.resnet50().to(device)
if arm == "ours":
    from dlmc_quant_b200 import quantize_model
    quantize_model(model, copy.deepcopy(Q.CFG), None)
x = torch.randn(128, 3, 224, 224, device=device)
t = torch.randint(0, 1000, (128,), device=device)
model.train()
with torch.no_grad():
    model(x[:8])
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
crit = nn.CrossEntropyLoss()


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), t)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
