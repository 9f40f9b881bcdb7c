"""TEST INFRASTRUCTURE ONLY - mint golden vectors from the UNMODIFIED reference.

Run in the authoring container (where /root/reference exists):

    python -m oracle.make_golden            # writes tests/golden/*.npz

Every case stores its inputs (small seeded tensors, plus adversarial special
values) and the outputs the reference's own code produced for them on CPU
(torch 2.11.0, fp32).  Module-level cases build the reference modules the way
`dlmc/utils/quantize.py:130-136` does (`__new__` + `__dict__.update` +
`initialize`) and capture the fake-quantized input/weight by intercepting
`_forward_func`, so the numbers are what `QBase.forward` / `RootQBase.forward` /
`FSPTQBase.forward` really computed - not a re-typed formula.

The fixtures travel to the GPU box; this script and the reference do not need to.
"""
import contextlib
import copy
import io
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 2333  # the reference's default seed (example/quantization/LSQ_config.yaml:8)


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


class Book:
    def __init__(self, name):
        self.name, self.arrays = name, {}

    def add(self, case, meta, inputs, outputs):
        assert "|" not in case
        self.arrays[f"{case}|meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        for k, v in inputs.items():
            self.arrays[f"{case}|in|{k}"] = _np(v)
        for k, v in outputs.items():
            self.arrays[f"{case}|out|{k}"] = _np(v)

    def save(self):
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        path = os.path.join(GOLDEN_DIR, self.name + ".npz")
        np.savez_compressed(path, **self.arrays)
        print(f"wrote {path}: {len(self.arrays)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


def special_values():
    """Ties, clamp boundaries, signed zeros, denormal, inf, nan, huge."""
    v = [0.0, -0.0, 0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 6.5, 7.0, 7.5, 8.5, 14.5, 15.0, 15.5, 16.0,
         -7.0, -7.5, -8.0, 1e-45, -1e-45, 1.17549435e-38, 3.4e38, -3.4e38, float("inf"),
         float("-inf"), float("nan"), 0.49999997, 0.50000006, 1.4999999, 1.5000001, 254.5, 255.5, 127.5, -127.5]
    return torch.tensor(v, dtype=torch.float32)


def gen_tensor(gen, shape, kind):
    if kind == "act":       # post-ReLU-like, ~50 % zeros
        return torch.relu(torch.randn(shape, generator=gen)) * 2
    if kind == "randn":
        return torch.randn(shape, generator=gen)
    if kind == "wt":
        return torch.randn(shape, generator=gen) * 0.02
    if kind == "shifted":   # asymmetric, strictly positive minimum
        return torch.rand(shape, generator=gen) * 3 + 0.75
    raise ValueError(kind)


# --------------------------------------------------------------------------- #
def golden_utils(ns, gen):
    b = Book("utils")
    u = ns.utils
    for signed, bits in [(True, 4), (False, 4), (True, 8), (False, 8)]:
        lo, hi = u.get_qrange(signed, bits)
        x = torch.cat([gen_tensor(gen, (509,), "randn") * 3, special_values()])
        for sname, scale, offset in [("s1", torch.tensor(1.0), torch.tensor(0)),
                                     ("sr", torch.tensor(0.3712), torch.tensor(-1.25)),
                                     ("s0", torch.tensor(0.0), torch.tensor(0.0))]:
            codes = u.quantize(x, scale, offset, lo, hi)
            y = u.emulate_quantize(x, scale, offset, lo, hi)
            b.add(f"a1_{'s' if signed else 'u'}{bits}_{sname}", {"lo": lo, "hi": hi, "signed": signed, "n_bits": bits},
                  {"x": x, "scale": scale, "offset": offset.float()}, {"codes": codes, "y": y})
        # per-channel broadcast [C,1]
        xc = gen_tensor(gen, (6, 37), "randn")
        sc = torch.rand(6, 1, generator=gen) * 0.2 + 0.01
        oc = torch.randn(6, 1, generator=gen) * 0.1
        b.add(f"a1_{'s' if signed else 'u'}{bits}_pc", {"lo": lo, "hi": hi},
              {"x": xc, "scale": sc, "offset": oc},
              {"codes": u.quantize(xc, sc, oc, lo, hi), "y": u.emulate_quantize(xc, sc, oc, lo, hi)})
    # grad_scale value: (s - s*g) + s*g is NOT always s
    s = torch.cat([torch.rand(4000, generator=gen) * 2 + 1e-3, torch.tensor([1.0, 0.0, 1e-30, 3e38])])
    for numel, qmax in [(3 * 32 * 32 * 128, 15), (2359296, 7), (802816 * 64, 15), (27, 127)]:
        g = 1 / np.sqrt(numel * qmax)
        b.add(f"grad_scale_{numel}_{qmax}", {"g": float(g)}, {"s": s}, {"value": u.grad_scale(s, float(g))})
    v = torch.cat([gen_tensor(gen, (300,), "randn") * 9, special_values()])
    b.add("round_pass", {}, {"v": v}, {"value": u.round_pass(v)})
    b.add("floor_pass", {}, {"v": v}, {"value": u.floor_pass(v)})
    b.save()


# --------------------------------------------------------------------------- #
def _build(new_type, base, qconfig):
    """dlmc/utils/quantize.py:130-133 - class swap without __init__."""
    m = new_type.__new__(new_type)
    m.__dict__.update(copy.deepcopy(base).__dict__)
    m.initialize(copy.deepcopy(qconfig))
    return m


def _capture(module):
    """Intercept _forward_func to see the fake-quantized (input, weight)."""
    seen = {}
    orig = module._forward_func

    def spy(inp, wt):
        inp.retain_grad() if inp.requires_grad else None
        wt.retain_grad() if wt.requires_grad else None
        seen["qx"], seen["qw"] = inp, wt
        return orig(inp, wt)

    module._forward_func = spy
    return seen


def _run_module(m, x, dy_seed, gen, train=True, steps=1):
    """Forward(+backward) `steps` times; return dict of tensors after the last step."""
    m.train(train)
    seen = _capture(m)
    out = {}
    for step in range(steps):
        for p in m.parameters():
            p.grad = None
        xin = x.detach().clone().requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            y = m(xin)
        if step == 0:
            out["dy"] = torch.randn(y.shape, generator=gen)
        if train:
            y.backward(out["dy"])
        out.update({"y": y.detach().clone(), "qx": seen["qx"].detach().clone(), "qw": seen["qw"].detach().clone()})
        if train:
            out["dx"] = xin.grad.clone()
            out["d_qx"] = seen["qx"].grad.clone()   # upstream grad of the fake-quantized input
            out["d_qw"] = seen["qw"].grad.clone()
            for n, p in m.named_parameters():
                if p.grad is not None:
                    out["grad_" + n] = p.grad.clone()
        for n, bf in m.named_buffers():
            if bf is not None and n != "org_weight":
                out["buf_" + n] = bf.detach().clone().float()
        for n, p in m.named_parameters():
            if n not in ("weight", "bias", "alpha"):
                out["param_" + n] = p.detach().clone()
    return out


def golden_qbase(ns, gen):
    b = Book("qbase")
    qnn = ns.modules
    combos = [
        ("mm_w4a4", {"type": "minmax_tensor", "args": {"n_bits": 4, "signed": True}},
         {"type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}}),
        ("mm_w8a8", {"type": "minmax_tensor", "args": {"n_bits": 8, "signed": True}},
         {"type": "minmax_tensor", "args": {"n_bits": 8, "signed": False}}),
        ("l2n_w4a4", {"type": "l2norm_tensor", "args": {"n_bits": 4, "signed": True}},
         {"type": "l2norm_tensor", "args": {"n_bits": 4, "signed": False}}),
        ("mm_w4a4_signed_act", {"type": "minmax_tensor", "args": {"n_bits": 4, "signed": True}},
         {"type": "minmax_tensor", "args": {"n_bits": 4, "signed": True}}),
    ]
    for name, wcfg, icfg in combos:
        qcfg = {"weight": dict(wcfg, enable=True), "input": dict(icfg, enable=True), "momentum": 0.1}
        conv = torch.nn.Conv2d(5, 7, 3, padding=1, bias=True)
        with torch.no_grad():
            conv.weight.copy_(gen_tensor(gen, conv.weight.shape, "wt"))
            conv.bias.copy_(gen_tensor(gen, conv.bias.shape, "wt"))
        x = gen_tensor(gen, (3, 5, 9, 9), "act" if not icfg["args"]["signed"] else "randn")
        m = _build(qnn.QConv2d, conv, qcfg)
        out = _run_module(m, x, 0, gen, train=True, steps=2)
        b.add(f"conv_{name}", {"qconfig": qcfg, "kind": "conv", "stride": 1, "padding": 1},
              {"x": x, "weight": conv.weight, "bias": conv.bias}, out)
        lin = torch.nn.Linear(33, 11)
        with torch.no_grad():
            lin.weight.copy_(gen_tensor(gen, lin.weight.shape, "wt"))
        xl = gen_tensor(gen, (8, 33), "act" if not icfg["args"]["signed"] else "randn")
        m = _build(qnn.QLinear, lin, qcfg)
        out = _run_module(m, xl, 0, gen, train=True, steps=1)
        b.add(f"linear_{name}", {"qconfig": qcfg, "kind": "linear"},
              {"x": xl, "weight": lin.weight, "bias": lin.bias}, out)
    # per-channel weight scale: only reachable by pre-shaping wt_scale
    qcfg = {"weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}},
            "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}}, "momentum": 0.1}
    conv = torch.nn.Conv2d(4, 6, 3, padding=1, bias=False)
    with torch.no_grad():
        conv.weight.copy_(gen_tensor(gen, conv.weight.shape, "wt") * torch.linspace(0.5, 3, 6).view(6, 1, 1, 1))
    x = gen_tensor(gen, (2, 4, 8, 8), "act")
    m = _build(qnn.QConv2d, conv, qcfg)
    m.wt_scale = torch.nn.Parameter(torch.ones(6, 1, 1, 1))
    out = _run_module(m, x, 0, gen, train=True, steps=1)
    b.add("conv_pcw_w4a4", {"qconfig": qcfg, "kind": "conv", "stride": 1, "padding": 1, "per_channel_weight": True},
          {"x": x, "weight": conv.weight}, out)
    b.save()


def golden_funlsq(ns, gen):
    b = Book("funlsq")
    F_ = ns.modules_function
    for name, signed, bits, kind in [("w4", True, 4, "wt"), ("a4", False, 4, "act"), ("w8", True, 8, "wt")]:
        lo, hi = ns.utils.get_qrange(signed, bits)
        w = gen_tensor(gen, (16, 45), kind).requires_grad_(True)
        scale = (w.detach().abs().max() / hi * 0.7).reshape(1).requires_grad_(True)
        offset = torch.zeros(1)
        g = float(1 / np.sqrt(w.numel() * hi))
        y = F_.FunLSQ.apply(w, scale, offset, lo, hi, g)
        dy = torch.randn(y.shape, generator=gen)
        y.backward(dy)
        b.add(f"funlsq_{name}", {"lo": lo, "hi": hi, "g": g},
              {"w": w, "scale": scale, "offset": offset, "dy": dy},
              {"y": y, "dw": w.grad, "dscale": scale.grad})
    b.save()


def golden_rootq(ns, gen):
    b = Book("rootq")
    RQ = ns.RootQ
    for name, wbits, abits, mom in [("w4a4", 4, 4, 0.1), ("w2a3", 2, 3, 0.25), ("w8a8", 8, 8, 0.1)]:
        qcfg = {"weight": {"enable": True, "type": None, "args": {"n_bits": wbits, "signed": False}},
                "input": {"enable": True, "type": None, "args": {"n_bits": abits, "signed": False}}, "momentum": mom}
        conv = torch.nn.Conv2d(5, 8, 3, padding=1, bias=False)
        with torch.no_grad():
            conv.weight.copy_(gen_tensor(gen, conv.weight.shape, "wt"))
        x = gen_tensor(gen, (4, 5, 10, 10), "act")
        m = _build(RQ.RootQConv2d, conv, qcfg)
        out1 = _run_module(m, x, 0, gen, train=True, steps=1)
        b.add(f"conv_{name}_step1", {"qconfig": qcfg, "kind": "conv", "stride": 1, "padding": 1, "steps": 1},
              {"x": x, "weight": conv.weight}, out1)
        # move the learnable bounds off their init so EMA/grad-mix matter, then step again
        m2 = _build(RQ.RootQConv2d, conv, qcfg)
        _run_module(m2, x, 0, gen, train=True, steps=1)
        with torch.no_grad():
            m2.wt_upper.mul_(0.83)
            m2.wt_lower.mul_(1.11)
            m2.in_scale.mul_(0.9)
            m2.wt_alpha.fill_(0.37)
        pre = {("pre_" + n): p.detach().clone() for n, p in m2.named_parameters() if n != "weight"}
        pre.update({("pre_" + n): bf.detach().clone() for n, bf in m2.named_buffers() if bf is not None})
        out2 = _run_module(m2, x, 0, gen, train=True, steps=1)
        b.add(f"conv_{name}_step2", {"qconfig": qcfg, "kind": "conv", "stride": 1, "padding": 1},
              dict({"x": x, "weight": conv.weight}, **pre), out2)
        out3 = _run_module(m2, x, 0, gen, train=False, steps=1)
        b.add(f"conv_{name}_eval", {"qconfig": qcfg, "kind": "conv", "stride": 1, "padding": 1},
              {"x": x, "weight": conv.weight}, out3)
    # free functions, special values
    fn = ns.rootq_function
    v = torch.cat([gen_tensor(gen, (200,), "randn") * 12, special_values()])
    b.add("clipping_0_15", {}, {"x": v, "upper": torch.tensor(15.0), "lower": torch.tensor(0.0)},
          {"value": fn.clipping(v, torch.tensor(15.0), 0)})
    b.add("clipping_w", {}, {"x": v * 0.01, "upper": torch.tensor(0.0731), "lower": torch.tensor(-0.0689)},
          {"value": fn.clipping(v * 0.01, torch.tensor(0.0731), torch.tensor(-0.0689))})
    b.save()


def golden_fsptq(ns, gen):
    """FSPTQBase.initialize hard-codes device='cuda' (FSPTQuant/base.py:47); redirect
    the allocation to CPU for the duration of initialize - reference code untouched."""
    b = Book("fsptq")
    FSP = ns.FSPTQuant
    real_zeros = torch.zeros

    def cpu_zeros(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)

    for name, wtype, itype, wbits, abits, recon in [
        ("mm_w8a8", "minmax_channel", "minmax_tensor", 8, 8, "none"),
        ("mm_w4a4", "minmax_channel", "minmax_tensor", 4, 4, "none"),
        ("l2_w4a8", "l2loss_channel", "l2loss_tensor", 4, 8, "none"),
        ("ada_w4a8", "minmax_channel", "minmax_tensor", 4, 8, "adaround"),
    ]:
        qcfg = {"weight": {"enable": True, "type": wtype, "recon_type": recon,
                           "args": {"n_bits": wbits, "signed": True, "ch_axis": 0}},
                "input": {"enable": True, "type": itype, "args": {"n_bits": abits, "signed": False}}, "momentum": 0.1}
        conv = torch.nn.Conv2d(4, 6, 3, padding=1, bias=True)
        with torch.no_grad():
            conv.weight.copy_(gen_tensor(gen, conv.weight.shape, "wt") * torch.linspace(0.5, 3, 6).view(6, 1, 1, 1))
        x = gen_tensor(gen, (3, 4, 8, 8), "act")
        torch.zeros = cpu_zeros
        try:
            m = _build(FSP.FSPTQConv2d, conv, qcfg)
        finally:
            torch.zeros = real_zeros
        out = _run_module(m, x, 0, gen, train=True, steps=1)
        if recon == "adaround":
            out["alpha"] = m.alpha.detach().clone()
            out["grad_alpha"] = m.alpha.grad.clone()
            oe = _run_module(m, x, 0, gen, train=False, steps=1)
            out["qw_eval"] = oe["qw"]
        b.add(f"conv_{name}", {"qconfig": qcfg, "kind": "conv", "stride": 1, "padding": 1},
              {"x": x, "weight": conv.weight, "bias": conv.bias}, out)
        lin = torch.nn.Linear(20, 9)
        xl = gen_tensor(gen, (5, 20), "act")
        torch.zeros = cpu_zeros
        try:
            m = _build(FSP.FSPTQLinear, lin, qcfg)
        finally:
            torch.zeros = real_zeros
        out = _run_module(m, xl, 0, gen, train=True, steps=1)
        b.add(f"linear_{name}", {"qconfig": qcfg, "kind": "linear"},
              {"x": xl, "weight": lin.weight, "bias": lin.bias}, out)
    b.save()


def golden_qgemm(ns, gen):
    """Whole-LAYER outputs of the reference's quantised Linear / 1x1 Conv2d modules in eval mode (what
    `_forward_func(q_input, q_weight)` returns, modules/base.py:140 and FSPTQuant/base.py:111-113), with the qparams
    the reference's own observers chose: the pin for the integer-code layer product (include/dlmcq.h, dlmcq_qgemm).
    K is a multiple of 16 everywhere so the same cases run through the TMA kernel on the GPU."""
    b = Book("qgemm")
    qnn, FSP = ns.modules, ns.FSPTQuant
    real_zeros = torch.zeros

    def cpu_zeros(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)

    def build(cls, base, qcfg, fsptq=False):
        if not fsptq:
            return _build(cls, base, qcfg)
        torch.zeros = cpu_zeros                   # FSPTQBase.initialize hard-codes device='cuda' (FSPTQuant/base.py:47)
        try:
            return _build(cls, base, qcfg)
        finally:
            torch.zeros = real_zeros

    def qcfg_of(wtype, itype, wbits, abits, act_signed=False, recon=None):
        w = {"enable": True, "type": wtype, "args": {"n_bits": wbits, "signed": True}}
        if "channel" in wtype:
            w["args"]["ch_axis"] = 0
        if recon is not None:
            w["recon_type"] = recon
        return {"weight": w, "input": {"enable": True, "type": itype, "args": {"n_bits": abits, "signed": act_signed}},
                "momentum": 0.1}

    def run(case, m, x, base, meta):
        out = _run_module(m, x, 0, gen, train=False, steps=1)
        b.add(case, meta, {"x": x, "weight": base.weight, "bias": base.bias if base.bias is not None else torch.zeros(0)}, out)

    # QBase family (modules/base.py): per-tensor weights as the reference allocates them
    for name, wbits, abits, kind, act_signed in [("w4a4", 4, 4, "act", False), ("w8a8_offset", 8, 8, "shifted", False),
                                                  ("w4a4_signed_act", 4, 4, "randn", True)]:
        qcfg = qcfg_of("minmax_tensor", "minmax_tensor", wbits, abits, act_signed)
        lin = torch.nn.Linear(64, 24)
        with torch.no_grad():
            lin.weight.copy_(gen_tensor(gen, lin.weight.shape, "wt"))
            lin.bias.copy_(gen_tensor(gen, lin.bias.shape, "randn"))
        run(f"qbase_linear_{name}", build(qnn.QLinear, lin, qcfg), gen_tensor(gen, (10, 64), kind), lin,
            {"qconfig": qcfg, "kind": "linear", "family": "qbase"})
        conv = torch.nn.Conv2d(32, 40, 1, bias=(name != "w4a4"))
        with torch.no_grad():
            conv.weight.copy_(gen_tensor(gen, conv.weight.shape, "wt"))
        run(f"qbase_conv1x1_{name}", build(qnn.QConv2d, conv, qcfg), gen_tensor(gen, (3, 32, 6, 5), kind), conv,
            {"qconfig": qcfg, "kind": "conv", "family": "qbase"})
    # QBase with a per-output-channel weight scale: only reachable by pre-shaping wt_scale
    qcfg = qcfg_of("minmax_channel", "minmax_tensor", 4, 4)
    conv = torch.nn.Conv2d(48, 36, 1, bias=True)
    with torch.no_grad():
        conv.weight.copy_(gen_tensor(gen, conv.weight.shape, "wt") * torch.linspace(0.5, 3, 36).view(36, 1, 1, 1))
    m = build(qnn.QConv2d, conv, qcfg)
    m.wt_scale = torch.nn.Parameter(torch.ones(36, 1, 1, 1))
    run("qbase_conv1x1_pcw_w4a4", m, gen_tensor(gen, (2, 48, 7, 7), "act"), conv,
        {"qconfig": qcfg, "kind": "conv", "family": "qbase", "per_channel_weight": True})
    # FSPTQ family (FSPTQuant/base.py): integer zero-point activations, symmetric per-channel weights (+1e-6)
    for name, wbits, abits in [("w8a8", 8, 8), ("w4a4", 4, 4)]:
        qcfg = qcfg_of("minmax_channel", "minmax_tensor", wbits, abits, recon="none")
        lin = torch.nn.Linear(48, 20)
        with torch.no_grad():
            lin.weight.copy_(gen_tensor(gen, lin.weight.shape, "wt") * torch.linspace(0.5, 2, 20).view(20, 1))
        run(f"fsptq_linear_{name}", build(FSP.FSPTQLinear, lin, qcfg, fsptq=True), gen_tensor(gen, (7, 48), "act"), lin,
            {"qconfig": qcfg, "kind": "linear", "family": "fsptq"})
        conv = torch.nn.Conv2d(16, 40, 1, bias=True)
        with torch.no_grad():
            conv.weight.copy_(gen_tensor(gen, conv.weight.shape, "wt") * torch.linspace(0.5, 3, 40).view(40, 1, 1, 1))
        run(f"fsptq_conv1x1_{name}", build(FSP.FSPTQConv2d, conv, qcfg, fsptq=True), gen_tensor(gen, (4, 16, 5, 5), "act"),
            conv, {"qconfig": qcfg, "kind": "conv", "family": "fsptq"})
    b.save()


class _Timeout(Exception):
    pass


def _bounded(fn, seconds=20):
    """Run fn() but give up after `seconds` (the reference's l2norm `while diff > eps`
    loops have no iteration cap and oscillate forever on some inputs)."""
    import signal

    def handler(signum, frame):
        raise _Timeout()

    old = signal.signal(signal.SIGALRM, handler)
    signal.alarm(seconds)
    try:
        return fn()
    except _Timeout:
        return None
    finally:
        signal.alarm(0)
        signal.signal(signal.SIGALRM, old)


def golden_observers(ns, gen):
    b = Book("observers")
    ops = ns.ops
    tensors = {
        "act4d": gen_tensor(gen, (4, 6, 7, 7), "act"),
        "randn4d": gen_tensor(gen, (4, 6, 5, 5), "randn"),
        "wt4d": gen_tensor(gen, (8, 5, 3, 3), "wt") * torch.linspace(0.3, 4, 8).view(8, 1, 1, 1),
        "shift2d": gen_tensor(gen, (6, 50), "shifted"),
        "wt2d": gen_tensor(gen, (10, 33), "wt"),
    }
    for tname, t in tensors.items():
        for signed, bits in [(True, 4), (False, 4), (True, 8), (False, 8)]:
            tag = f"{tname}_{'s' if signed else 'u'}{bits}"
            meta = {"n_bits": bits, "signed": signed}
            s, o = ops.quantize_minmax_tensor(t, bits, signed)
            b.add(f"minmax_tensor_{tag}", meta, {"t": t}, {"scale": s, "offset": o.float()})
            for ax in ([0, 1] if t.dim() == 4 else [0]):
                s, o = ops.quantize_minmax_channel(t, bits, signed, ch_axis=ax)
                b.add(f"minmax_channel{ax}_{tag}", dict(meta, ch_axis=ax), {"t": t}, {"scale": s, "offset": o})
            if t.dim() == 4:
                s, o = ops.quantize_minmax_pixel(t, bits, signed)
                b.add(f"minmax_pixel_{tag}", meta, {"t": t}, {"scale": s, "offset": o})
            with contextlib.redirect_stdout(io.StringIO()) as cap:
                s, o = ops.quantize_l2loss_tensor(t, bits, signed)
            picked = int(cap.getvalue().split()[0]) if cap.getvalue().strip() else -1
            b.add(f"l2loss_tensor_{tag}", dict(meta, picked=picked), {"t": t}, {"scale": s, "offset": o.float()})
            s, o = ops.quantize_l2loss_channel(t.clone(), bits, signed, ch_axis=0)
            b.add(f"l2loss_channel_{tag}", dict(meta, ch_axis=0), {"t": t}, {"scale": s, "offset": o})
            if tname in ("randn4d", "wt4d", "wt2d", "shift2d") or not signed:
                r = _bounded(lambda: ops.quantize_l2norm_tensor(t, bits, signed))
                if r is None:
                    print(f"  l2norm_tensor_{tag}: reference does not terminate, skipped")
                else:
                    b.add(f"l2norm_tensor_{tag}", meta, {"t": t}, {"scale": r[0], "offset": r[1].float()})
                r = _bounded(lambda: ops.quantize_l2norm_channel(t, bits, signed, ch_axis=0))
                if r is None:
                    print(f"  l2norm_channel_{tag}: reference does not terminate, skipped")
                else:
                    b.add(f"l2norm_channel_{tag}", dict(meta, ch_axis=0), {"t": t}, {"scale": r[0], "offset": r[1]})
    b.save()


def _randomize_bn(bn, gen, tiny_var=False):
    c = bn.num_features
    with torch.no_grad():
        bn.weight.copy_(torch.randn(c, generator=gen))            # negative gammas included
        bn.bias.copy_(torch.randn(c, generator=gen) * 0.5)
        bn.running_mean.copy_(torch.randn(c, generator=gen) * 0.3)
        bn.running_var.copy_(torch.rand(c, generator=gen) * 2 + 0.01)
        if tiny_var:
            bn.running_var[0] = 1e-8                                # exercises the +1e-7 / +eps
            bn.running_var[1] = 0.0
    return bn


def _bn_pack(bn):
    return dict(gamma=bn.weight.detach().clone(), beta=bn.bias.detach().clone(),
                mean=bn.running_mean.detach().clone(), var=bn.running_var.detach().clone())


def golden_reparam(gen):
    """merge_bn (dlmc/utils/merge_bn.py:45-113) and RepVGGBlock.switch_to_deploy (repvgg.py:92-147) run by the
    reference's own code; the folded kernels also go through the reference's per-channel min/max observer."""
    rp = ref_shim.load_reparam()
    ns = ref_shim.load()
    b = Book("reparam")
    nn = torch.nn
    # ---- merge_bn: "0"/"1" naming (case 1) and conv1/bn1 naming (case 2), bias / no bias / groups
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 8, 3, bias=True)
            self.bn1 = nn.BatchNorm2d(8)
            self.conv2 = nn.Conv2d(8, 12, 3, groups=2, bias=False)
            self.bn2 = nn.BatchNorm2d(12)
            self.block = nn.Sequential(nn.Conv2d(12, 6, 1, bias=False), nn.BatchNorm2d(6))
    torch.manual_seed(SEED)
    net = Net()
    pairs = [("conv1", "bn1"), ("conv2", "bn2"), ("block.0", "block.1")]
    from operator import attrgetter
    for k, (cn, bnn) in enumerate(pairs):
        _randomize_bn(attrgetter(bnn)(net), gen, tiny_var=(k == 0))
    with torch.no_grad():
        net.conv1.weight[0, 0, 0, 0] = -0.0
        net.conv2.weight[3].zero_()
    inputs = {}
    for cn, bnn in pairs:
        conv, bn = attrgetter(cn)(net), attrgetter(bnn)(net)
        inputs[cn] = dict(w=conv.weight.detach().clone(),
                          bias=None if conv.bias is None else conv.bias.detach().clone(), **_bn_pack(bn))
    with contextlib.redirect_stdout(io.StringIO()):
        merged = rp.merge_bn.merge_bn(net)           # inplace=False modifies `net` itself (merge_bn.py:61-62)
    assert merged is net
    for cn, bnn in pairs:
        conv = attrgetter(cn)(net)
        assert isinstance(attrgetter(bnn)(net), nn.Identity)
        inp = {k: v for k, v in inputs[cn].items() if v is not None}
        so, oo = ns.ops.quantize_minmax_channel(conv.weight.detach(), n_bits=8, signed=True, ch_axis=0)
        b.add(f"merge_bn.{cn}", {"has_bias": inputs[cn]["bias"] is not None}, inp,
              {"w": conv.weight.detach(), "bias": conv.bias.detach(), "obs_scale": so, "obs_offset": oo})
    # ---- RepVGG blocks: identity / no identity / grouped
    for name, (cin, cout, stride, groups) in {"id": (8, 8, 1, 1), "noid": (4, 8, 2, 1), "grouped": (8, 8, 1, 2)}.items():
        with contextlib.redirect_stdout(io.StringIO()):
            blk = rp.repvgg.RepVGGBlock(cin, cout, 3, stride=stride, padding=1, groups=groups)
        _randomize_bn(blk.rbr_dense.bn, gen, tiny_var=(name == "id"))
        _randomize_bn(blk.rbr_1x1.bn, gen)
        if blk.rbr_identity is not None:
            _randomize_bn(blk.rbr_identity, gen)
        with torch.no_grad():
            blk.rbr_dense.conv.weight[1, 0, 1, 1] = -0.0
            blk.rbr_dense.conv.weight[2, 1, 0, 0] = -0.0
        inp = {"k3": blk.rbr_dense.conv.weight.detach().clone(), "k1": blk.rbr_1x1.conv.weight.detach().clone()}
        for tag, bn in (("bn3", blk.rbr_dense.bn), ("bn1", blk.rbr_1x1.bn), ("bnid", blk.rbr_identity)):
            if bn is not None:
                for k, v in _bn_pack(bn).items():
                    inp[f"{tag}_{k}"] = v
        eps = blk.rbr_dense.bn.eps
        blk.eval()
        blk.switch_to_deploy()
        w, bias = blk.rbr_reparam.weight.detach(), blk.rbr_reparam.bias.detach()
        so, oo = ns.ops.quantize_minmax_channel(w, n_bits=8, signed=True, ch_axis=0)
        b.add(f"repvgg.{name}", {"groups": groups, "eps": eps, "has_id": "bnid_gamma" in inp}, inp,
              {"w": w, "bias": bias, "obs_scale": so, "obs_offset": oo})
    b.save()


def main():
    ns = ref_shim.load()
    torch.manual_seed(SEED)
    torch.set_num_threads(1)     # reduction order of the stored sums is then machine-independent
    if sys.argv[1:] == ["qgemm"]:                # this book has its own generator: mint it without touching the others
        golden_qgemm(ns, torch.Generator().manual_seed(SEED + 11))
        return
    gen = torch.Generator().manual_seed(SEED)
    golden_utils(ns, gen)
    golden_qbase(ns, gen)
    golden_funlsq(ns, gen)
    golden_rootq(ns, gen)
    golden_fsptq(ns, gen)
    golden_observers(ns, gen)
    golden_reparam(torch.Generator().manual_seed(SEED + 7))
    golden_qgemm(ns, torch.Generator().manual_seed(SEED + 11))


if __name__ == "__main__":
    main()
