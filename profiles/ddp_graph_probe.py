"""Whole-step CUDA-graph capture UNDER DistributedDataParallel (the step is host-bound there: profiles/
r02_ddp_gpu_busy_n2.jsonl): ResNet-50 channels_last batch 128 per GPU, fused QAT arm and the un-quantised model, eager
vs graphed.  PyTorch's documented recipe: DDP constructed inside a side stream, >= 11 eager warm-up iterations on that
stream (DDP's bucket rebuild and the reducer's first-iteration logic must have happened), then capture.
    torchrun --nproc-per-node N profiles/ddp_graph_probe.py"""
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
import torchvision  # noqa: E402

os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")      # required for NCCL work inside CUDA graphs
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True


def run(arm, graphed, steps=20):
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
    crit = nn.CrossEntropyLoss()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        if world > 1:
            model = nn.parallel.DistributedDataParallel(model, device_ids=[dev.index])
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=5e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = crit(model(x), t)
            loss.backward()
            opt.step()
            return loss
        for _ in range(12):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if graphed:
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            loss = step()
        run_step = g.replay
    else:
        run_step = step
    for _ in range(3):
        run_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        run_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / steps * 1e3
    del model, opt
    torch.cuda.empty_cache()
    return round(dt, 2)


CASES = (("ours_fused", True),) if os.environ.get("DLMCQ_PROBE_TRACE") else (("fp32", False), ("fp32", True), ("ours_fused", False), ("ours_fused", True))
for arm, graphed in CASES:
    try:
        ms = run(arm, graphed)
        res = {"ms_per_step": ms, "images_per_s": round(128 * world / ms * 1e3, 1)}
    except Exception as e:
        import traceback
        if rank == 0 and os.environ.get("DLMCQ_PROBE_TRACE"):
            traceback.print_exc()
        res = {"error": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0:
        print(json.dumps({"arm": arm, "graphed": graphed, "n_gpus": world, **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
