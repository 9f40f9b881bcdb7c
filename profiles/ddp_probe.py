"""Why does the fused QAT arm lose more than the un-quantised model when DistributedDataParallel wraps it?
ms/step of ResNet-50 (channels_last, batch 128 per GPU) under DDP variants, plus the HOST time of a step (the loop
timed without waiting for the GPU): if host time ~ step time the step is host-bound, not all-reduce-bound.
    torchrun --nproc-per-node 2 profiles/ddp_probe.py"""
import copy
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True


def run(arm, ddp_kwargs, n_groups=None, steps=10, batch=128):
    import torchvision
    torch.manual_seed(2333)
    model = torchvision.models.resnet50().to(dev).to(memory_format=torch.channels_last)
    if arm != "fp32":
        from dlmc_quant_b200 import quantize_model
        quantize_model(model, copy.deepcopy(bench.QAT_CFG), None)
    x = torch.randn(batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 1000, (batch,), device=dev)
    model.train()
    with torch.no_grad():
        model(x[:8])
    if arm != "fp32":
        from dlmc_quant_b200.quantize import group_weight_quantizers
        if n_groups != 0:
            group_weight_quantizers(model, n_groups=n_groups)
        if arm == "ours_fused":
            from dlmc_quant_b200.fuse import fuse_bn_act_quant
            fuse_bn_act_quant(model)
    if world > 1 and ddp_kwargs is not None:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], **ddp_kwargs)
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
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    host = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / steps * 1e3
    del model, opt
    torch.cuda.empty_cache()
    return round(total, 2), round(host, 2)


cases = [("fp32", {}, None), ("fp32", None, None), ("ours_fused", None, None), ("ours_fused", {}, None),
         ("ours_fused", {"broadcast_buffers": False}, None), ("ours_fused", {"gradient_as_bucket_view": True}, None),
         ("ours_fused", {"static_graph": True}, None), ("ours_fused", {}, 1), ("ours_fused", {}, 16), ("ours_fused", {}, 0),
         ("ours_fused", {"bucket_cap_mb": 100}, None)]
for arm, kw, ng in cases:
    try:
        total, host = run(arm, kw, ng)
    except Exception as e:
        total, host = f"{type(e).__name__}: {e}"[:120], None
    if rank == 0:
        print(json.dumps({"arm": arm, "ddp": "none (local step)" if kw is None else kw, "weight_groups": ng,
                          "ms_per_step": total, "host_ms_per_step": host}), flush=True)
if world > 1:
    dist.destroy_process_group()
