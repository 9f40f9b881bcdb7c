"""End-to-end ResNet-50 W4A4 QAT images/s on synthetic ImageNet-shaped batches (BASELINE.json configs[1],
methodology of the reference's example/benchmark/benchmark.py:168-197: 2 warm-up steps, wall clock,
images / elapsed; SGD nesterov lr 0.01 as in benchmark.yaml:44-50).

    python profiles/qat_images_per_s.py [--batch 64] [--steps 20] [--arm ours|eager|fp32|all]
    torchrun --nproc-per-node N profiles/qat_images_per_s.py ...      (DDP, per-GPU batch fixed)

Arms:  ours  - torchvision resnet50 with every Conv2d/Linear swapped by dlmc_quant_b200.quantize_model
               (per-tensor A4 activations, per-channel W4 weights, fused kernels)
       eager - the same quantizers as the reference computes them: the eager torch op chain of
               modules/base.py:96-102,131-133 differentiated by autograd, on the same GPU
               ("PyTorch eager on B200").  Test infrastructure, uses oracle/restate.py.
       fp32  - the un-quantised model, for scale.
       bf16 / ours_bf16 - (opt-in: --arm bf16,ours_bf16) the same two models under torch.autocast(bfloat16):
               bf16 activations go through the bf16 kernels (fp32 arithmetic, one RNE rounding on store,
               fp32 scale gradients); an extension beyond the fp32-only reference, reported separately.
The convolutions are cuDNN library calls in all arms; only the fake-quant path differs."""
import argparse
import copy
import json
import math
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = {"weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}},
       "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}},
       "exclude_layers": [], "override_options": [], "momentum": 0.1}


def make_eager(model):
    """Swap in modules that run the reference's eager chain (oracle restatement) on the GPU."""
    from oracle import restate as R

    class EagerQ(nn.Module):
        def __init__(self, base):
            super().__init__()
            self.base = base
            self.in_scale = nn.Parameter(torch.ones(1, device=base.weight.device))
            shape = [base.weight.shape[0]] + [1] * (base.weight.dim() - 1)
            self.wt_scale = nn.Parameter(torch.ones(shape, device=base.weight.device))
            self.ready = False

        def forward(self, x):
            w = self.base.weight
            if not self.ready:
                s, o = R.obs_minmax_tensor(x.detach(), 4, False)
                self.in_scale.data.copy_(s.reshape(1))
                self.in_offset = o.detach()
                ws, wo = R.obs_minmax_channel(w.detach(), 4, True, ch_axis=0)
                self.wt_scale.data.copy_(ws)
                self.wt_offset = wo.detach()
                self.ready = True
            x = R.fq_affine(x, self.in_scale, self.in_offset, 0, 15, 1 / math.sqrt(x.numel() * 15))
            wq = R.fq_affine(w, self.wt_scale, self.wt_offset, -7, 7, 1 / math.sqrt(w.numel() * 7))
            b = self.base
            if isinstance(b, nn.Conv2d):
                return F.conv2d(x, wq, b.bias, b.stride, b.padding, b.dilation, b.groups)
            return F.linear(x, wq, b.bias)

    for name, m in list(model.named_modules()):
        for cname, c in list(m.named_children()):
            if isinstance(c, (nn.Conv2d, nn.Linear)):
                setattr(m, cname, EagerQ(c))
    return model


def run(arm, args, device, world):
    import torchvision
    torch.manual_seed(2333)
    model = torchvision.models.resnet50().to(device)
    amp = arm.endswith("bf16")
    if arm in ("ours", "ours_bf16"):
        from dlmc_quant_b200 import quantize_model
        quantize_model(model, copy.deepcopy(CFG), None)
    elif arm == "eager":
        make_eager(model)
    x = torch.randn(args.batch, 3, 224, 224, device=device)
    if args.channels_last:
        model = model.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 1000, (args.batch,), device=device)
    model.train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        model(x[:8])            # lazy observer init before DDP wraps the parameters (shapes may change)
    if arm.startswith("ours") and not args.per_layer_weights:
        from dlmc_quant_b200.quantize import group_weight_quantizers
        group_weight_quantizers(model)      # all 54 weight tensors: one launch per direction instead of 54 x 3
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[device.index])
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
    crit = nn.CrossEntropyLoss()

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            loss = crit(model(x), t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    float(loss)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    return args.batch * world * args.steps / dt, dt / args.steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--arm", default="all")
    ap.add_argument("--per-layer-weights", action="store_true", help="ours: no step-level weight grouping")
    ap.add_argument("--channels-last", action="store_true",
                    help="model and input in channels_last memory format (all arms); the fake-quant kernels index "
                         "dense channels_last tensors directly")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    torch.backends.cudnn.benchmark = True        # example/benchmark/benchmark.py:203-205 (benchmark.yaml:14 cudnn: true)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    out = {"model": "torchvision resnet50, W4 per-channel / A4 per-tensor QAT, fp32 (TF32 convs: torch default)",
           "per_gpu_batch": args.batch, "n_gpus": world, "steps": args.steps, "data": "synthetic 3x224x224",
           "memory_format": "channels_last" if args.channels_last else "contiguous (NCHW)"}
    for arm in (["fp32", "eager", "ours"] if args.arm == "all" else args.arm.split(",")):
        ips, ms = run(arm, args, device, world)
        out[arm] = {"images_per_s": round(ips, 1), "ms_per_step": round(ms, 2)}
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
