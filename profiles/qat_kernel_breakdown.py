"""Where the QAT step's GPU time goes: per-kernel CUDA time of 3 steps of ResNet-50 W4A4 at batch 128 (torch.profiler).

    python profiles/qat_kernel_breakdown.py [fp32|ours|ours_fused] [nchw|channels_last] [batch]

Prints a table of GPU kernels grouped by name (total time, calls, average) plus the step's wall time."""
import copy
import os
import sys
import time
from collections import defaultdict

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = {"weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}},
       "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}},
       "exclude_layers": [], "override_options": [], "momentum": 0.1}
arm = sys.argv[1] if len(sys.argv) > 1 else "ours_fused"
fmt = sys.argv[2] if len(sys.argv) > 2 else "channels_last"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 128
import torchvision
torch.backends.cudnn.benchmark = True
torch.manual_seed(2333)
device = torch.device("cuda", 0)
model = torchvision.models.resnet50().to(device)
if arm.startswith("ours"):
    from dlmc_quant_b200 import quantize_model
    quantize_model(model, copy.deepcopy(CFG), None)
x = torch.randn(batch, 3, 224, 224, device=device)
if fmt == "channels_last":
    model = model.to(memory_format=torch.channels_last)
    x = x.contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 1000, (batch,), device=device)
model.train()
with torch.no_grad():
    model(x[:8])
if arm.startswith("ours"):
    from dlmc_quant_b200.quantize import group_weight_quantizers
    group_weight_quantizers(model)
if arm == "ours_fused":
    from dlmc_quant_b200.fuse import fuse_bn_act_quant
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
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:110]][0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        agg[e.name[:110]][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"arm={arm} format={fmt} batch={batch}: wall {wall:.2f} ms/step, GPU kernel time {tot / 3e3:.2f} ms/step")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{us / 3e3:8.3f} ms/step {100 * us / tot:5.1f}%  calls/step {n / 3:6.1f}  avg {us / n:8.1f} us  {name}")
