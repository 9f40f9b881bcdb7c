"""SURVEY.md 8f row f3 on BASELINE.json configs[2] shapes: RepVGG-A0 (22 blocks, 8.3 M deploy-form weights) from
train form to deploy form (repvgg.py:92-147) + per-channel W8 min/max qparams (ops.py:121-140).

GPU arm: `dlmcq_fold_grouped` (one launch for all blocks, statistics fused) + one finalisation launch.
CPU arm: the oracle port of the reference's eager chain (get_equivalent_kernel_bias per block, then
quantize_minmax_channel re-reading the fused kernel) on the host cores.  One JSON line.

    python profiles/reparam_c3.py [--reps 50]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def repvgg_a0_blocks():
    """(cin, cout, has_identity) of RepVGG-A0's 22 blocks (repvgg.py:150-200: num_blocks [2,4,14,1],
    width multipliers [0.75, 0.75, 0.75, 2.5])."""
    blocks, cin = [(3, 48, False)], 48
    for planes, n in ((48, 2), (96, 4), (192, 14), (1280, 1)):
        for i in range(n):
            blocks.append((cin, planes, i > 0 and cin == planes))
            cin = planes
    assert len(blocks) == 22
    return blocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    from dlmc_quant_b200 import reparam as P
    from oracle import restate as R
    gen = torch.Generator().manual_seed(2333)
    eps = 1e-5

    def bn(c):
        return (torch.randn(c, generator=gen), torch.randn(c, generator=gen) * 0.5, torch.randn(c, generator=gen) * 0.3,
                torch.rand(c, generator=gen) * 2 + 0.01, eps)
    cpu = []
    for cin, cout, has_id in repvgg_a0_blocks():
        cpu.append(dict(k3=torch.randn(cout, cin, 3, 3, generator=gen) * 0.05, k1=torch.randn(cout, cin, 1, 1, generator=gen) * 0.05,
                        bn3=bn(cout), bn1=bn(cout), bnid=bn(cin) if has_id else None))
    dev = lambda t: tuple(v.cuda() if isinstance(v, torch.Tensor) else v for v in t)
    entries = []
    for e in cpu:
        w = e["k3"].cuda()
        entries.append(dict(mode="repvgg", w=w, w_out=torch.empty_like(w), bias_out=torch.empty(w.shape[0], device="cuda"),
                            w1=e["k1"].cuda(), bn=dev(e["bn3"]), bn1=dev(e["bn1"]),
                            bn_id=dev(e["bnid"]) if e["bnid"] is not None else None))
    chans = [e["w"].shape[0] for e in entries]

    def gpu_pass():
        stats = P.fold_grouped(entries)
        return P.observe_folded(stats, chans, 8, True)

    qp = gpu_pass()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        gpu_pass()
    b.record()
    torch.cuda.synchronize()
    gpu_s = a.elapsed_time(b) * 1e-3 / args.reps

    torch.set_num_threads(os.cpu_count())

    def cpu_pass():
        out = []
        for e in cpu:
            w, bias = R.repvgg_fuse(e["k3"], e["bn3"], e["k1"], e["bn1"], e["bnid"], 1)
            out.append((w, bias) + R.obs_minmax_channel(w, 8, True, ch_axis=0))
        return out
    ref = cpu_pass()
    t0 = time.perf_counter()
    for _ in range(3):
        cpu_pass()
    cpu_s = (time.perf_counter() - t0) / 3
    exact = all(torch.equal(e["w_out"].cpu(), r[0]) and torch.equal(e["bias_out"].cpu(), r[1]) and
                torch.equal(s.cpu().reshape(-1), r[2].reshape(-1)) for e, r, (s, _) in zip(entries, ref, qp))
    n = sum(e["k3"].numel() for e in cpu)
    print(json.dumps({"config": "C3 RepVGG-A0 train form -> deploy form + per-channel W8 min/max qparams (row f3)",
                      "blocks": len(cpu), "weight_elements": n, "bit_exact_vs_cpu_port": exact,
                      "gpu": {"us": round(gpu_s * 1e6, 1), "launches": 2, "includes": "descriptor upload (host) + fold + finalise"},
                      "cpu_reference_port": {"us": round(cpu_s * 1e6, 1), "cores": os.cpu_count()},
                      "speedup": round(cpu_s / gpu_s, 1)}))


if __name__ == "__main__":
    main()
