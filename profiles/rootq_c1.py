"""BASELINE.json configs[0] ("C1"): ResNet-18 W4A4 RootQ fake-quant fwd+bwd on a CIFAR-shaped batch
(model/classification/cifarresnet_large.py:132, batch 128x3x32x32, RootQ_cifar10_config.yaml:21): all 21
quantised layers' activation + weight quantizers, forward and root-estimator backward.

GPU arm: the RootQ kernels through the functional API, replayed as one CUDA graph.  CPU arm: the oracle
port of RootQBase.forward (eager torch + autograd) on the host cores - the reference's own path for this
config.  One JSON line.   python profiles/rootq_c1.py [--batch 128] [--steps 50]"""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cifar_resnet18_layers():
    """(activation C,H,W per image, weight shape) of the 21 quantised layers; SURVEY.md App. B totals."""
    L = [((3, 32, 32), (64, 3, 3, 3))]
    cin, hw = 64, 32
    for planes, stride in [(64, 1), (128, 2), (256, 2), (512, 2)]:
        for b in range(2):
            s = stride if b == 0 else 1
            L.append(((cin, hw, hw), (planes, cin, 3, 3)))
            L.append(((planes, hw // s, hw // s), (planes, planes, 3, 3)))
            if b == 0 and (s != 1 or cin != planes):
                L.append(((cin, hw, hw), (planes, cin, 1, 1)))
            cin, hw = planes, hw // s
    L.append(((512,), (10, 512)))
    assert len(L) == 21 and sum(math.prod(a) for a, _ in L) == 667136 and sum(math.prod(w) for _, w in L) == 11164352
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cpu-passes", type=int, default=3)
    ap.add_argument("--per-layer", action="store_true", help="six launches per layer instead of the grouped launches")
    ap.add_argument("--act-per-layer", action="store_true",
                    help="grouped prepare / weights, but one activation launch per layer and direction (round 1)")
    args = ap.parse_args()
    from dlmc_quant_b200 import functional as F
    from oracle import restate as R
    layers = cifar_resnet18_layers()
    gen = torch.Generator().manual_seed(2333)
    mom, bits = 0.1, 4
    lo, hi = 0, 2 ** bits - 1
    cpu, dev = [], []
    for i, (a, w) in enumerate(layers):
        x = torch.randn((args.batch,) + a, generator=gen)
        if i:
            x = torch.relu(x) * 2
        wt = torch.randn(w, generator=gen) * 0.03
        dyx, dyw = torch.randn(x.shape, generator=gen), torch.randn(wt.shape, generator=gen)
        in_scale = R.rootq_act_init(x, lo, hi)
        up, dn = R.rootq_wt_init(wt, hi)
        cpu.append((x, wt, dyx, dyw, in_scale, up.float(), dn.float()))
        c = lambda t: t.cuda()
        dev.append(dict(x=c(x), w=c(wt), dyx=c(dyx), dyw=c(dyw), in_scale=c(in_scale), run=c(in_scale).clone(),
                        up=c(up.float()), dn=c(dn.float()), alpha=torch.tensor(0.25).cuda(), rup=c(up.float()).clone(),
                        rdn=c(dn.float()).clone(), g_i=1 / math.sqrt(x.numel() * hi), g_w=1 / math.sqrt(wt.numel() * hi)))
    elems = sum(d["x"].numel() + d["w"].numel() for d in dev)

    def step_per_layer():
        for d in dev:
            sa = F.rootq_act_prepare(d["in_scale"], d["run"], mom, d["g_i"], lo, hi, True)
            sw = F.rootq_wt_prepare(d["up"], d["dn"], d["alpha"], d["rup"], d["rdn"], mom, d["g_w"], lo, hi, True)
            d["y"] = F.rootq_act_forward(d["x"], sa)
            d["wq"] = F.rootq_wt_forward(d["w"], sw)
            d["sa"], d["sw"] = sa, sw
        for d in reversed(dev):
            d["dx"], d["ds"] = F.rootq_act_backward(d["x"], d["dyx"], d["sa"])
            d["dw"], d["gw"] = F.rootq_wt_backward(d["w"], d["dyw"], d["sw"])

    # grouped: one prepare launch for the 42 quantizers, one forward and one backward (+ finalise) launch for the
    # 21 weight tensors; the activations stay one launch per layer and direction (they arrive layer by layer)
    grp = F.GroupedRootQ("cuda")
    quantizers = []
    for d in dev:
        quantizers.append(dict(kind="act", in_scale=d["in_scale"], run_scale=d["run"], momentum=mom, g=d["g_i"], lo=lo,
                               hi=hi, training=True))
        quantizers.append(dict(kind="wt", upper=d["up"], lower=d["dn"], alpha=d["alpha"], run_upper=d["rup"],
                               run_lower=d["rdn"], momentum=mom, g=d["g_w"], lo=lo, hi=hi, training=True))
    for d in dev:
        d["wq"], d["dw"], d["gw"] = torch.empty_like(d["w"]), torch.empty_like(d["w"]), torch.empty(3, device="cuda")

    def step_grouped():
        states = grp.prepare(quantizers)
        grp.wt_forward([dict(w=d["w"], out=d["wq"], state=states[2 * i + 1]) for i, d in enumerate(dev)])
        for i, d in enumerate(dev):
            d["y"] = F.rootq_act_forward(d["x"], states[2 * i])
        for i, d in reversed(list(enumerate(dev))):
            d["dx"], d["ds"] = F.rootq_act_backward(d["x"], d["dyx"], states[2 * i])
        grp.wt_backward([dict(w=d["w"], dy=d["dyw"], out=d["dw"], state=states[2 * i + 1], grads=d["gw"])
                         for i, d in enumerate(dev)])

    for d in dev:
        d["y"], d["dx"], d["ds"] = torch.empty_like(d["x"]), torch.empty_like(d["x"]), torch.empty(1, device="cuda")

    def step_all_grouped():
        # the quantizer set has every tensor resident at once: the 21 activation tensors go through ONE launch per
        # direction as well (dlmcq_rootq_act_forward_grouped / _backward_grouped) - 8 launches for the whole step
        states = grp.prepare(quantizers)
        grp.wt_forward([dict(w=d["w"], out=d["wq"], state=states[2 * i + 1]) for i, d in enumerate(dev)])
        grp.act_forward([dict(w=d["x"], out=d["y"], state=states[2 * i]) for i, d in enumerate(dev)])
        grp.act_backward([dict(w=d["x"], dy=d["dyx"], out=d["dx"], state=states[2 * i], grads=d["ds"])
                          for i, d in enumerate(dev)])
        grp.wt_backward([dict(w=d["w"], dy=d["dyw"], out=d["dw"], state=states[2 * i + 1], grads=d["gw"])
                         for i, d in enumerate(dev)])

    step = step_per_layer if args.per_layer else (step_grouped if args.act_per_layer else step_all_grouped)
    launches = 21 * 6 if args.per_layer else (21 * 2 + 4 if args.act_per_layer else 8)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps

    torch.set_num_threads(os.cpu_count())
    best = float("inf")
    for _ in range(args.cpu_passes):
        t0 = time.perf_counter()
        for x, wt, dyx, dyw, in_scale, up, dn in cpu:
            R.rootq_act_fwd_bwd(x, in_scale, in_scale.clone(), mom, lo, hi, dyx)
            R.rootq_wt_fwd_bwd(wt, up, dn, torch.tensor(0.25), up.clone(), dn.clone(), mom, lo, hi, dyw)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({
        "config": "C1: cifar ResNet-18 W4A4 RootQ, batch %d x 3x32x32, 21 layers, act+weight quantizers fwd+bwd" % args.batch,
        "elements_per_step": elems, "launches_per_step": launches,
        "gpu": {"ms_per_step": round(ms, 4), "gbps_algorithmic_20B": round(20 * elems / ms / 1e6, 1),
                "gelem_s": round(elems / ms / 1e6, 2)},
        "cpu_reference_port": {"ms_per_step": round(best * 1e3, 1), "gbps_algorithmic_20B": round(20 * elems / best / 1e9, 3),
                               "cores": os.cpu_count(), "kind": "port (oracle restatement of RootQBase.forward + autograd)"},
        "speedup": round(best * 1e3 / ms, 1)}), flush=True)


if __name__ == "__main__":
    main()
