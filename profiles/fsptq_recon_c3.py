"""Row f1, measured: FSPTQ / RepAPQ calibration + block reconstruction of a RepVGG-A0-shaped deploy network (BASELINE
configs[2]: W8A8, min/max observers, 1 024 calibration images in batches of 128) - wall time of
  ours   dlmc_quant_b200.recon.FSPTQReconstructor: full-precision outputs of ALL blocks cached in one pass, the quantised
         pass for block k stops at block k, caches stay in HBM;
  port   the reference procedure restated literally (trainer/fsptq_trainer.py:28-112): for EVERY block both models run
         over the whole calibration set, the hooked tensors go to the CPU, are concatenated there and come back.
Both arms use this package's FSPTQ modules and the same number of fitting iterations per block (the reference's default
is 20 000; --iters bounds the run), so the difference is the data path of the procedure, not the kernels."""
import argparse
import copy
import json
import os
import sys
import time

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlmc_quant_b200 import quantize_model  # noqa: E402
from dlmc_quant_b200.recon import FSPTQReconstructor, l2_loss  # noqa: E402
from dlmc_quant_b200.scalar.FSPTQuant.base import FSPTQBase  # noqa: E402

CFG = {"weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 8, "signed": True, "ch_axis": 0}},
       "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 8, "signed": False}},
       "exclude_layers": [], "override_options": [], "momentum": 0.1}


class Block(nn.Module):
    """A deploy-mode RepVGG block: one 3x3 convolution + ReLU (model/classification/repvgg.py after switch_to_deploy)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.rbr_reparam = nn.Conv2d(cin, cout, 3, stride, 1, bias=True)

    def forward(self, x):
        return torch.relu(self.rbr_reparam(x))


class RepVGGA0Deploy(nn.Module):
    """RepVGG-A0 widths / depths: stage0 48, stages [2, 4, 14, 1] x [48, 96, 192, 1280] (22 blocks), linear 1000."""

    def __init__(self):
        super().__init__()
        layers, cin = [Block(3, 48, 2)], 48
        for n, w in zip([2, 4, 14, 1], [48, 96, 192, 1280]):
            for i in range(n):
                layers.append(Block(cin, w, 2 if i == 0 else 1))
                cin = w
        self.stages = nn.Sequential(*layers)
        self.linear = nn.Linear(1280, 1000)

    def forward(self, x):
        return self.linear(self.stages(x).mean((2, 3)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=128)
    args = ap.parse_args()
    torch.manual_seed(2333)
    torch.backends.cudnn.benchmark = True
    fp = RepVGGA0Deploy().cuda().eval()
    batches = [torch.randn(args.batch, 3, 224, 224, device="cuda") for _ in range(args.images // args.batch)]

    def build():
        m = copy.deepcopy(fp)
        quantize_model(m, copy.deepcopy(CFG), None, quantization_type="FSPTQ")
        return m

    out = {"model": "RepVGG-A0 deploy shapes (22 blocks + linear), FSPTQ W8A8, minmax observers",
           "calibration": f"{args.images} synthetic 3x224x224 images in batches of {args.batch}",
           "iterations_per_block": args.iters}
    ours = build()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = FSPTQReconstructor(ours, fp, block_types=(Block,), epochs=args.iters)
    rec.run(batches, generator=torch.Generator().manual_seed(1))
    torch.cuda.synchronize()
    out["ours_s"] = round(time.perf_counter() - t0, 2)

    naive = build()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = torch.Generator().manual_seed(1)
    for (name, module), fp_module in zip(naive.named_modules(), fp.modules()):
        if not ((isinstance(module, FSPTQBase) and name in ["conv1", "linear"]) or type(module) is Block):
            continue
        cin, cout = [], []
        h1 = fp_module.register_forward_hook(lambda m, i, o: cout.append(o.cpu()))
        h2 = module.register_forward_hook(lambda m, i, o: cin.append(i[0].cpu()))
        naive.eval()
        for data in batches:
            with torch.no_grad():
                fp(data)
                naive(data)
        block_input, block_output = torch.cat(cin).cuda(), torch.cat(cout).cuda()
        h1.remove()
        h2.remove()
        opt, sched = rec.generate_optimizer(module)
        naive.train()
        for i in range(args.iters):
            idx = torch.randperm(block_input.size(0), generator=g)[:64].cuda()
            opt.zero_grad()
            loss = l2_loss(block_output[idx], module(block_input[idx]))
            loss.backward()
            opt.step()
            sched.step()
    torch.cuda.synchronize()
    out["port_s"] = round(time.perf_counter() - t0, 2)
    out["speedup"] = round(out["port_s"] / out["ours_s"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
