"""Weight-space re-parameterisation feeding the per-channel observers.

    merge_bn(model, ...)               dlmc/utils/merge_bn.py:45-113   BatchNorm2d folded into its Conv2d
    repvgg_model_convert(model, ...)   model/classification/repvgg.py:125-147, 297-305   RepVGG blocks -> 3x3 convs

Same call signatures and module surgery as the reference (BN replaced by Identity; `rbr_reparam` created,
branches deleted, `deploy = True`).  What changes: every layer of the model goes through ONE launch of
`dlmcq_fold_grouped` (the reference runs ~10 eager ops per layer), and the same pass leaves each folded
row's {min, max, max|w|, sum|w|}, so `observe_folded` turns them into per-channel (scale, offset) -
`quantize_minmax_channel`, ops.py:121-140 - without reading the folded weights again.  Bit-identical to the
reference's eager arithmetic (tests/test_reparam.py, fixtures minted from the reference's own code).

Reference quirk kept: `merge_bn(inplace=True)` deep-copies the model and `inplace=False` (the default)
modifies the argument - the flag is inverted in merge_bn.py:61-62.  The `bitmixer_func` option is not
offered (the BitMixer package is absent from the reference tree)."""
import copy
import ctypes as C
from operator import attrgetter

import torch
from torch import nn

from . import _lib
from . import functional as F
from .quantize import attrsetter, get_layers

__all__ = ["DEFAULT_BN_MAPPING_FN", "DEFAULT_CONV_MAPPING_FN", "merge_bn", "repvgg_model_convert", "fold_grouped",
           "observe_folded"]


def DEFAULT_CONV_MAPPING_FN(bn_name):
    """merge_bn.py:13-26: `layer1.conv1.1 -> layer1.conv1.0`, `layer1.bn1 -> layer1.conv1`."""
    *parent, base = bn_name.split(".")
    if base.isdecimal():
        return ".".join(parent + [str(int(base) - 1)])
    if "bn" in base:
        return ".".join(parent + [base.replace("bn", "conv")])
    return None


def DEFAULT_BN_MAPPING_FN(conv_name):
    """merge_bn.py:29-42 (the reference forgets its `return`; the intended inverse mapping is returned here)."""
    *parent, base = conv_name.split(".")
    if base.isdecimal():
        return ".".join(parent + [str(int(base) + 1)])
    if "conv" in base:
        return ".".join(parent + [base.replace("conv", "bn")])
    return None


def _f32(t, name):
    if not t.is_cuda:
        raise _lib.DlmcqError(f"{name} must live on a CUDA device: the fold runs in libdlmcq.so, there is no CPU path")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous fp32 tensor")
    return t.data_ptr()


def fold_grouped(entries):
    """One launch for all `entries` (dicts).  Common keys: mode ('merge_bn' | 'repvgg'), w [C, ...], w_out,
    bias_out [C], bn = (gamma, beta, mean, var[, eps]); merge_bn: bias or None; repvgg: w1 [C, cin_g, 1, 1],
    bn1, bn_id or None.  Returns the [sum C, 4] statistics of the folded rows, in entry order."""
    if not entries:
        return None
    device = entries[0]["w"].device
    arr = (_lib.FoldItem * len(entries))()
    chans = [0]
    for e in entries:
        chans.append(chans[-1] + e["w"].shape[0])
    stats = torch.empty(chans[-1], 4, dtype=torch.float32, device=device)
    for i, e in enumerate(entries):
        it, w = arr[i], e["w"]
        c = w.shape[0]
        it.w, it.w_out, it.bias_out = _f32(w, "w"), _f32(e["w_out"], "w_out"), _f32(e["bias_out"], "bias_out")
        it.channels, it.inner = c, w.numel() // c
        it.stats = stats.data_ptr() + 16 * chans[i]
        g, b, m, v = e["bn"][:4]
        it.gamma, it.beta, it.mean, it.var = _f32(g, "gamma"), _f32(b, "beta"), _f32(m, "mean"), _f32(v, "var")
        if e["mode"] == "merge_bn":
            it.mode = _lib.FOLD_MERGE_BN
            it.bias = _f32(e["bias"], "bias") if e.get("bias") is not None else None
        else:
            it.mode = _lib.FOLD_REPVGG
            if w.dim() != 4 or w.shape[2] != 3 or w.shape[3] != 3:
                raise ValueError("RepVGG fold expects a [C, cin/g, 3, 3] kernel")     # repvgg.py:29
            it.cin_g, it.ksize, it.eps = w.shape[1], 3, float(e["bn"][4])
            it.w1 = _f32(e["w1"], "w1")
            g1, b1, m1, v1, eps1 = e["bn1"]
            it.gamma1, it.beta1, it.mean1, it.var1 = _f32(g1, "gamma1"), _f32(b1, "beta1"), _f32(m1, "mean1"), _f32(v1, "var1")
            it.eps1 = float(eps1)
            if e.get("bn_id") is not None:
                gi, bi, mi, vi, epsi = e["bn_id"]
                it.gamma_id, it.beta_id, it.mean_id, it.var_id = (_f32(gi, "gamma_id"), _f32(bi, "beta_id"),
                                                                  _f32(mi, "mean_id"), _f32(vi, "var_id"))
                it.eps_id = float(epsi)
    items = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    prefix = torch.tensor(chans, dtype=torch.int64).to(device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().dlmcq_fold_grouped(items.data_ptr(), prefix.data_ptr(), len(entries), chans[-1],
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return stats


def observe_folded(stats, channels, n_bits, signed, allow_offset=True):
    """ops.py:121-140 on the statistics `fold_grouped` left behind: one finalisation launch for all layers;
    returns [(scale [C,1,1,1], offset [C,1,1,1]), ...] split by `channels`."""
    scale, offset = F.minmax_from_stats(stats, n_bits, signed, allow_offset)
    out, at = [], 0
    for c in channels:
        out.append((scale[at:at + c].reshape(-1, 1, 1, 1), offset[at:at + c].reshape(-1, 1, 1, 1)))
        at += c
    return out


def _bn(bn, with_eps=False):
    t = (bn.weight.data, bn.bias.data, bn.running_mean.data, bn.running_var.data)
    return t + (bn.eps,) if with_eps else t


def merge_bn(model, mapping_fn=DEFAULT_CONV_MAPPING_FN, inplace=False, allow_missing=False, return_stats=False):
    """merge_bn.py:45-113.  With return_stats=True also returns {conv name: [C, 4] statistics of the folded
    weight} for `observe_folded`."""
    if inplace:                                   # sic - merge_bn.py:61-62
        model = copy.deepcopy(model)
    all_layers = get_layers(model, filter_types=(nn.Conv2d, nn.BatchNorm2d))
    entries, names, surgery = [], [], []
    for layer in all_layers:
        module = attrgetter(layer)(model)
        if not isinstance(module, nn.BatchNorm2d):
            continue
        map_name = mapping_fn(layer)
        if map_name is None or map_name not in all_layers:
            msg = f"[MergeBN] Could not find Conv2d that match {layer}"
            if not allow_missing:
                raise ValueError(msg)
            print(msg)
            continue
        conv = attrgetter(map_name)(model)
        weight = conv.weight.data
        had_bias = conv.bias is not None
        if not had_bias:                          # merge_bn.py:92-93
            conv.bias = nn.Parameter(torch.zeros(weight.shape[0], device=weight.device))
        entries.append(dict(mode="merge_bn", w=weight, w_out=weight, bias=conv.bias.data if had_bias else None,
                            bias_out=conv.bias.data, bn=_bn(module)))
        names.append(map_name)
        surgery.append(layer)
    stats = fold_grouped(entries)
    for layer in surgery:                         # merge_bn.py:101-110
        *parent, base = layer.split(".")
        parent_mod = attrgetter(".".join(parent))(model) if parent else model
        attrsetter(base)(parent_mod, nn.Identity())
    if not return_stats:
        return model
    out, at = {}, 0
    for name, e in zip(names, entries):
        c = e["w"].shape[0]
        out[name] = stats[at:at + c]
        at += c
    return model, out


def repvgg_model_convert(model, save_path=None, do_copy=True, return_stats=False):
    """repvgg.py:297-305 + RepVGGBlock.switch_to_deploy (:125-147) for every block, in one launch."""
    if do_copy:
        model = copy.deepcopy(model)
    blocks = [(n, m) for n, m in model.named_modules()
              if hasattr(m, "rbr_dense") and hasattr(m, "rbr_1x1") and not hasattr(m, "rbr_reparam")]
    entries = []
    for _, blk in blocks:
        conv = blk.rbr_dense.conv
        rep = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, stride=conv.stride, padding=conv.padding,
                        dilation=conv.dilation, groups=conv.groups, bias=True).to(conv.weight.device)
        ident = getattr(blk, "rbr_identity", None)
        entries.append(dict(mode="repvgg", w=conv.weight.data.contiguous(), w_out=rep.weight.data, bias_out=rep.bias.data,
                            bn=_bn(blk.rbr_dense.bn, True), w1=blk.rbr_1x1.conv.weight.data.contiguous(),
                            bn1=_bn(blk.rbr_1x1.bn, True), bn_id=_bn(ident, True) if ident is not None else None,
                            rep=rep))
    stats = fold_grouped(entries)
    for (_, blk), e in zip(blocks, entries):
        blk.rbr_reparam = e["rep"]
        del blk.rbr_dense
        del blk.rbr_1x1
        if hasattr(blk, "rbr_identity"):
            del blk.rbr_identity
        if hasattr(blk, "id_tensor"):
            del blk.id_tensor
        blk.deploy = True
    if save_path is not None:
        torch.save(model.state_dict(), save_path)
    if not return_stats:
        return model
    out, at = {}, 0
    for (name, _), e in zip(blocks, entries):
        c = e["w"].shape[0]
        out[name] = stats[at:at + c]
        at += c
    return model, out
