"""Is the DDP QAT step GPU-bound or host-bound?  torch.profiler over 5 steps: wall time per step against the summed
duration of the GPU kernels per CUDA stream (run with DLMCQ_NO_PDL=1 so that kernel durations are not inflated by
programmatic-dependent-launch waits)."""
import copy
import json
import os
import sys
import time
from collections import defaultdict

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torchvision  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True
for arm in sys.argv[1:] or ["fp32", "ours_fused"]:
    torch.manual_seed(2333)
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
        fuse_bn_act_quant(model)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[dev.index])
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
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 5 * 1e3
    per_stream = defaultdict(float)
    nccl = 0.0
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            d = e.device_time if hasattr(e, "device_time") else e.cuda_time
            if "nccl" in e.name.lower():
                nccl += d
            else:
                per_stream["compute"] += d
    if rank == 0:
        print(json.dumps({"arm": arm, "n_gpus": world, "wall_ms_per_step_under_profiler": round(wall, 2),
                          "compute_kernels_ms_per_step": round(per_stream["compute"] / 5e3, 2),
                          "nccl_kernels_ms_per_step": round(nccl / 5e3, 2)}), flush=True)
    del model, opt
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
