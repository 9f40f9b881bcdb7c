"""BASELINE.json configs[2] and [3]: post-training calibration observer pass.

  C3  RepVGG-A0 (deploy form, model/classification/repvgg.py:201,297) RepAPQ/FSPTQ W8A8:
      activations min/max + 80-point MSE clip sweep per tensor (batch 128, FSPTQ_config.yaml:18,26),
      weights min/max + MSE sweep per output channel, then the FSPTQ fake-quant forward of both.
  C4  MobileOne-S0 (deploy shapes; the model is not in the reference tree - hand-listed from the
      architecture: depthwise 3x3 + pointwise 1x1 pairs, widths 48/48/128/256/1024, depths 2/8/10/1)
      W4A8 with per-channel weight quantizers (row lengths down to 9).

GPU arm: observer kernels through dlmc_quant_b200.scalar.ops (the reference's API names).  CPU arm: the
oracle port of the same observers on the host cores; the per-channel sweep is a Python double loop in the
reference (ops.py:175-194), so it is timed on a bounded sample of channels and scaled (flagged).
One JSON line per config.   python profiles/calibration_c3_c4.py [--batch 128]"""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def repvgg_a0_layers():
    L = [((3, 224, 224), (48, 3, 3, 3))]
    cin, hw = 48, 112
    for planes, blocks in [(48, 2), (96, 4), (192, 14), (1280, 1)]:
        for b in range(blocks):
            L.append(((cin, hw, hw), (planes, cin, 3, 3)))
            if b == 0:
                hw //= 2
            cin = planes
    L.append(((1280,), (1000, 1280)))
    assert len(L) == 23 and sum(math.prod(a) for a, _ in L) == 1882880 and sum(math.prod(w) for _, w in L) == 8303888
    return L


def mobileone_s0_layers():
    L = [((3, 224, 224), (48, 3, 3, 3))]
    cin, hw = 48, 112
    for planes, blocks in [(48, 2), (128, 8), (256, 10), (1024, 1)]:
        for b in range(blocks):
            L.append(((cin, hw, hw), (cin, 1, 3, 3)))          # depthwise 3x3 (stride 2 in the first block)
            if b == 0:
                hw //= 2
            L.append(((cin, hw, hw), (planes, cin, 1, 1)))     # pointwise 1x1
            cin = planes
    L.append(((1024,), (1000, 1024)))
    return L


def run_config(name, layers, batch, wbits, abits, cpu_act_batch, cpu_rows):
    from dlmc_quant_b200 import functional as F
    from dlmc_quant_b200._lib import FORM_SYM, FORM_ZP
    from dlmc_quant_b200.scalar import ops
    from oracle import restate as R
    gen = torch.Generator().manual_seed(2333)
    data = []
    for i, (a, w) in enumerate(layers):
        x = torch.randn((batch,) + a, generator=gen)
        if i:
            x = torch.relu(x) * 2
        wt = torch.randn(w, generator=gen) * 0.03
        data.append((x, wt))
    dev = [(x.cuda(), w.cuda()) for x, w in data]
    act_elems = sum(x.numel() for x, _ in data)
    wt_elems = sum(w.numel() for _, w in data)
    rows = sum(w.shape[0] for _, w in data)
    wlo, whi = -(2 ** (wbits - 1) - 1), 2 ** (wbits - 1) - 1

    def gpu_pass(sweep):
        out = []
        for x, w in dev:
            if sweep:
                s_in, o_in = ops.quantize_l2loss_tensor(x, abits, False)
                s_w, _ = ops.quantize_l2loss_channel(w, wbits, True, ch_axis=0)
            else:
                s_in, o_in = ops.quantize_minmax_tensor(x, abits, False)
                s_w, _ = ops.quantize_minmax_channel(w, wbits, True, ch_axis=0)
            qx = F.fq_forward(x, s_in, o_in, 0, 2 ** abits - 1, FORM_ZP)
            qw = F.fq_forward(w, s_w.reshape(-1) + 1e-6, None, wlo, whi, FORM_SYM, ch_axis=0)
            out.append((qx, qw))
        return out

    res = {}
    for sweep in (False, True):
        gpu_pass(sweep)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            gpu_pass(sweep)
        torch.cuda.synchronize()
        res["mse_sweep" if sweep else "minmax"] = (time.perf_counter() - t0) / 3

    # CPU reference port: min/max pass in full; sweeps on a bounded sample, scaled linearly
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    for x, w in data:
        s_in, o_in = R.obs_minmax_tensor(x, abits, False)
        s_w, _ = R.obs_minmax_channel(w, wbits, True, ch_axis=0)
        R.fq_zp(x, s_in, o_in, 0, 2 ** abits - 1)
        R.fq_sym(w, s_w + 1e-6, wlo, whi)
    cpu_minmax = time.perf_counter() - t0
    t0 = time.perf_counter()
    n_act = 0
    for x, _ in data:
        xs = x[:cpu_act_batch]
        R.obs_l2loss_tensor(xs, abits, False)
        n_act += xs.numel()
    cpu_act_sweep = (time.perf_counter() - t0) * act_elems / n_act
    t0 = time.perf_counter()
    n_rows = 0
    for _, w in data[:: max(1, len(data) // 6)]:
        ws = w[:cpu_rows]
        R.obs_l2loss_channel(ws.clone(), wbits, True)
        n_rows += ws.shape[0]
    cpu_row_sweep = (time.perf_counter() - t0) * rows / n_rows
    cpu_sweep = cpu_minmax + cpu_act_sweep + cpu_row_sweep
    print(json.dumps({
        "config": name, "batch": batch, "layers": len(layers), "activation_elements": act_elems,
        "weight_elements": wt_elems, "weight_channels": rows,
        "gpu_s": {"minmax_observers+fake_quant_fwd": round(res["minmax"], 5),
                  "mse_sweep_observers+fake_quant_fwd": round(res["mse_sweep"], 5)},
        "cpu_reference_port_s": {"minmax_observers+fake_quant_fwd": round(cpu_minmax, 3),
                                 "mse_sweep_observers+fake_quant_fwd (scaled from sample)": round(cpu_sweep, 1),
                                 "cores": os.cpu_count(),
                                 "sample": f"activation sweep on batch {cpu_act_batch} of {batch}, channel sweep on "
                                           f"{n_rows} of {rows} rows, both scaled linearly"},
        "speedup": {"minmax": round(cpu_minmax / res["minmax"], 1), "mse_sweep": round(cpu_sweep / res["mse_sweep"], 1)}}),
        flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    args = ap.parse_args()
    run_config("C3 RepVGG-A0 deploy, FSPTQ W8A8 calibration", repvgg_a0_layers(), args.batch, 8, 8, 4, 8)
    run_config("C4 MobileOne-S0 deploy shapes, W4A8 per-channel PTQ calibration", mobileone_s0_layers(), args.batch, 4, 8, 4, 8)


if __name__ == "__main__":
    main()
