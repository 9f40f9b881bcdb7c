"""Thin torch-facing wrappers over the C ABI (include/dlmcq.h).

PyTorch is plumbing here: device memory, the current stream, autograd glue.  Every function
extracts raw device pointers and the current CUDA stream and calls libdlmcq.so; nothing in
this module computes on the CPU or falls back to eager torch ops."""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import (BF16, F32, FORM_A1, FORM_AFFINE, FORM_SYM, FORM_ZP, ROOTQ_STATE_FLOATS, STATS_PER_CHANNEL,
                   SWEEP_CANDIDATES, DlmcqError, Layout, QParams)

__all__ = ["fq_forward", "fq_backward", "dequantize", "obs_stats", "minmax_from_stats", "absmean_from_stats",
           "sweep_tensor", "sweep_channel", "sweep_channel_grouped", "kth_values", "l2norm_fixed_point", "adaround_forward", "adaround_backward",
           "adaround_init_alpha", "rootq_act_prepare", "rootq_act_forward", "rootq_act_backward", "rootq_wt_prepare",
           "rootq_wt_forward", "rootq_wt_backward", "GroupedFakeQuant", "GroupedRootQ", "HostFakeQuant", "layout_of"]


def _require_cuda(t, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise DlmcqError(f"{name} must be a CUDA tensor: this path has no CPU implementation")


def _dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise DlmcqError(f"unsupported dtype {t.dtype}: fp32 and bf16 only")


def _raw_stream(index=None):
    """Current CUDA stream handle as an int, without building a torch.cuda.Stream object (per-layer hot path)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device() if index is None else index)


def _stream_ptr():
    return C.c_void_p(_raw_stream())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_layouts = {}


def layout_of(t, ch_axis=None):
    """[outer, channels, inner] view of a contiguous tensor (see include/dlmcq.h).  Cached per
    (shape, axis, dtype): the structs are read-only for the library."""
    key = (t.shape, ch_axis, t.dtype)
    lay = _layouts.get(key)
    if lay is not None:
        return lay
    if ch_axis is None:
        lay = Layout(1, 1, t.numel(), _dtype_code(t))
    else:
        shape = list(t.shape)
        ax = ch_axis % len(shape)
        lay = Layout(math.prod(shape[:ax]), shape[ax], math.prod(shape[ax + 1:]), _dtype_code(t))
    if len(_layouts) < 4096:
        _layouts[key] = lay
    return lay


def _qvec(v, channels, device, name):
    """scale / offset as a float32 device vector of `channels` entries (no sync when already on device)."""
    if v is None:
        return None
    if (isinstance(v, torch.Tensor) and v.dtype is torch.float32 and v.device == device and v.numel() == channels
            and v.is_contiguous()):
        return v                                   # parameters / buffers already in kernel form: only data_ptr is read
    if not isinstance(v, torch.Tensor):
        v = torch.full((channels,), float(v), dtype=torch.float32, device=device)
    if v.device != device or v.dtype != torch.float32:
        v = v.to(device=device, dtype=torch.float32)
    v = v.detach().reshape(-1)
    if v.numel() == 1 and channels != 1:
        v = v.expand(channels)
    if v.numel() != channels:
        raise DlmcqError(f"{name} has {v.numel()} entries, layout has {channels} channels")
    return v.contiguous()


_workspaces = {}


def _workspace(device, nbytes):
    """Zero-initialised scratch, one per (device, stream); kernels leave it zeroed."""
    key = (device.index, _raw_stream(device.index))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_scratches = {}


def _scratch(device, nbytes):
    """Plain scratch memory (no zero contract), one buffer per (device, stream), grown on demand."""
    key = (device.index, _raw_stream(device.index))
    buf = _scratches.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _scratches[key] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
    return buf


_ws_bytes = {}


def _ws_for(t, lay):
    key = (lay.outer, lay.channels, lay.inner)
    n = _ws_bytes.get(key)
    if n is None:
        n = _ws_bytes[key] = _lib.lib().dlmcq_workspace_bytes(C.byref(lay))
    return _workspace(t.device, n), n


class _NoCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NOCTX = _NoCtx()


def _on(device):
    """Device guard for a launch; free when `device` already is the current device (the per-layer hot path)."""
    return _NOCTX if device.index == torch.cuda.current_device() else torch.cuda.device(device)


# --------------------------------------------------------------------------------------
def dense_as_is(x, ch_axis):
    """True when the kernels can index x's storage directly: contiguous, or a dense channels_last 4-D tensor
    quantised per tensor / per dim-0 channel (dim 0 stays outermost, and within one tensor or one row the
    element order does not matter) - so channels_last models are served without a layout copy."""
    return x.is_contiguous() or (x.dim() == 4 and (ch_axis is None or ch_axis == 0) and
                                 x.is_contiguous(memory_format=torch.channels_last))


def _dense(x, ch_axis):
    return x if dense_as_is(x, ch_axis) else x.contiguous()


def fq_forward(x, scale, offset, lo, hi, form, g=0.0, ch_axis=None, want_codes=False, want_y=True):
    """Fused fake-quant forward -> y (and/or the integer codes as a float tensor), in x's memory format."""
    _require_cuda(x, "x")
    x = _dense(x.detach(), ch_axis)
    lay = layout_of(x, ch_axis)
    s = _qvec(scale, lay.channels, x.device, "scale")
    o = _qvec(offset, lay.channels, x.device, "offset")
    y = torch.empty_like(x) if want_y else None
    codes = torch.empty_like(x) if want_codes else None
    qp = QParams(form, int(lo), int(hi), float(g), s.data_ptr(), o.data_ptr() if o is not None else None)
    with _on(x.device):
        _lib.check(_lib.lib().dlmcq_fq_forward(_ptr(x), _ptr(y), _ptr(codes), C.byref(lay), C.byref(qp), _stream_ptr()))
    if want_y and want_codes:
        return y, codes
    return y if want_y else codes


def fq_backward(x, dy, scale, offset, lo, hi, form, g=0.0, ch_axis=None, want_doffset=False):
    """Fused backward: one pass over (x, dy) -> dx, dscale[channels] (, doffset[channels])."""
    _require_cuda(x, "x")
    _require_cuda(dy, "dy")
    x = _dense(x.detach(), ch_axis)
    dy = dy.detach()
    if dy.dtype != x.dtype or dy.shape != x.shape:
        raise DlmcqError("dy must match x in dtype and shape")
    if dy.stride() != x.stride() or not dense_as_is(dy, ch_axis):      # same element order as x
        dy = dy.contiguous() if x.is_contiguous() else dy.contiguous(memory_format=torch.channels_last)
    lay = layout_of(x, ch_axis)
    s = _qvec(scale, lay.channels, x.device, "scale")
    o = _qvec(offset, lay.channels, x.device, "offset")
    dx = torch.empty_like(x)
    ds = torch.empty(lay.channels, dtype=torch.float32, device=x.device)
    do = torch.empty(lay.channels, dtype=torch.float32, device=x.device) if want_doffset else None
    qp = QParams(form, int(lo), int(hi), float(g), s.data_ptr(), o.data_ptr() if o is not None else None)
    with _on(x.device):
        ws, n = _ws_for(x, lay)
        _lib.check(_lib.lib().dlmcq_fq_backward(_ptr(x), _ptr(dy), _ptr(dx), _ptr(ds), _ptr(do), C.byref(lay),
                                                C.byref(qp), _ptr(ws), n, _stream_ptr()))
    return (dx, ds, do) if want_doffset else (dx, ds)


def dequantize(codes, scale, offset, ch_axis=None):
    _require_cuda(codes, "codes")
    codes = codes.detach().contiguous()
    lay = layout_of(codes, ch_axis)
    s = _qvec(scale, lay.channels, codes.device, "scale")
    o = _qvec(offset, lay.channels, codes.device, "offset")
    y = torch.empty_like(codes)
    with torch.cuda.device(codes.device):
        _lib.check(_lib.lib().dlmcq_dequantize(_ptr(codes), _ptr(y), C.byref(lay), _ptr(s), _ptr(o), _stream_ptr()))
    return y


def export_codes(x, scale, offset, lo, hi, form, g=0.0, ch_axis=None, pack4=False):
    """Integer codes as bytes: int8 (signed ranges) / uint8, or two 4-bit codes per byte (`pack4`)."""
    _require_cuda(x, "x")
    x = x.detach().contiguous()
    lay = layout_of(x, ch_axis)
    s = _qvec(scale, lay.channels, x.device, "scale")
    o = _qvec(offset, lay.channels, x.device, "offset")
    n = x.numel()
    out = torch.empty((n + 1) // 2 if pack4 else n, dtype=torch.uint8 if (pack4 or lo >= 0) else torch.int8, device=x.device)
    qp = QParams(form, int(lo), int(hi), float(g), s.data_ptr(), o.data_ptr() if o is not None else None)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dlmcq_export_codes(_ptr(x), _ptr(out), C.byref(lay), C.byref(qp), int(bool(pack4)),
                                                 _stream_ptr()))
    return out if pack4 else out.reshape(x.shape)


def import_codes(codes, shape, scale, offset, lo, hi, form, g=0.0, ch_axis=None, pack4=False, dtype=torch.float32):
    """Unpack + dequantise exported codes; bit-identical to fq_forward's output for the same qparams."""
    _require_cuda(codes, "codes")
    y = torch.empty(shape, dtype=dtype, device=codes.device)
    lay = layout_of(y, ch_axis)
    s = _qvec(scale, lay.channels, y.device, "scale")
    o = _qvec(offset, lay.channels, y.device, "offset")
    qp = QParams(form, int(lo), int(hi), float(g), s.data_ptr(), o.data_ptr() if o is not None else None)
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().dlmcq_import_codes(_ptr(codes.contiguous()), _ptr(y), C.byref(lay), C.byref(qp),
                                                 int(bool(pack4)), _stream_ptr()))
    return y


def ste_value(x, mode):
    """mode 0 round_pass value, 1 floor_pass value, 2 sgn (utils.py:29-37, RootQ/function.py:5-8)."""
    _require_cuda(x, "x")
    x = x.detach().contiguous()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dlmcq_ste_value(_ptr(x), _ptr(y), x.numel(), _dtype_code(x), int(mode), _stream_ptr()))
    return y


def grad_scale_value(s, g):
    """utils.py:24-27 value (s - s*g) + s*g for a float32 device tensor of scales."""
    _require_cuda(s, "scale")
    s = s.detach().contiguous()
    if s.dtype != torch.float32:
        raise DlmcqError("scales are float32")
    out = torch.empty_like(s)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().dlmcq_grad_scale_value(_ptr(s), _ptr(out), s.numel(), float(g), _stream_ptr()))
    return out


# --------------------------------------------------------------------------------------
# observers
def obs_stats(x, ch_axis=None, abs_input=False):
    """One read of x -> float32 [channels, 4] = (min, max, max|x|, sum|x|)."""
    _require_cuda(x, "x")
    x = x.detach().contiguous()
    lay = layout_of(x, ch_axis)
    stats = torch.empty(lay.channels, STATS_PER_CHANNEL, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        ws, n = _ws_for(x, lay)
        _lib.check(_lib.lib().dlmcq_obs_stats(_ptr(x), _ptr(stats), C.byref(lay), int(bool(abs_input)), _ptr(ws), n,
                                              _stream_ptr()))
    return stats


SCALAR_DIVISION = "ieee"


def set_scalar_division(mode):
    """How the min/max observers evaluate `range / (2**n - 1)` (ops.py:23,32,126,135, a python-scalar divisor):
    "ieee"        true float32 division - the reference on the CPU, what the committed fixtures pin (default);
    "cuda_eager"  multiply by the float32 reciprocal - what eager PyTorch computes for the same expression ON CUDA
                  (ATen's true-division kernel special-cases CPU-scalar divisors), i.e. the reference as its users
                  run it on a GPU; up to 1 ulp away from "ieee".  Returns the previous mode."""
    global SCALAR_DIVISION
    if mode not in ("ieee", "cuda_eager"):
        raise DlmcqError("scalar division mode is 'ieee' or 'cuda_eager'")
    prev, SCALAR_DIVISION = SCALAR_DIVISION, mode
    return prev


def minmax_from_stats(stats, n_bits, signed, allow_offset=True, division=None):
    ch = stats.shape[0]
    scale = torch.empty(ch, dtype=torch.float32, device=stats.device)
    offset = torch.empty(ch, dtype=torch.float32, device=stats.device)
    mode = _lib.DIV_CUDA_EAGER if (division or SCALAR_DIVISION) == "cuda_eager" else _lib.DIV_IEEE
    with torch.cuda.device(stats.device):
        _lib.check(_lib.lib().dlmcq_obs_minmax_finalize_mode(_ptr(stats), _ptr(scale), _ptr(offset), ch, int(n_bits),
                                                             int(bool(signed)), int(bool(allow_offset)), mode,
                                                             _stream_ptr()))
    return scale, offset


def absmean_from_stats(stats, count, mul_a, mul_b, mode):
    ch = stats.shape[0]
    out = torch.empty(ch, dtype=torch.float32, device=stats.device)
    with torch.cuda.device(stats.device):
        _lib.check(_lib.lib().dlmcq_obs_absmean_finalize(_ptr(stats), _ptr(out), ch, float(count), float(mul_a),
                                                         float(mul_b), int(mode), _stream_ptr()))
    return out


def kth_values(x, ranks, abs_input=False, reduce_hist=None, fast=True):
    """Exact order statistics: the ranks[j]-th smallest elements (1-based, at most two ranks) of x or |x| by a
    3-pass radix select - equal to torch.kthvalue.  reduce_hist (multi-GPU): called on the int32 histogram of each
    pass (e.g. an all-reduce SUM) so that the result is the order statistic of the union of all ranks' tensors."""
    _require_cuda(x, "x")
    x = x.detach().contiguous()
    ranks = [int(r) for r in ranks]
    if not 1 <= len(ranks) <= 2 or min(ranks) < 1:
        raise DlmcqError("kth_values takes one or two ranks >= 1")
    h = _lib.lib()
    values = torch.empty(2, dtype=torch.float32, device=x.device)
    flags = 1 if abs_input else 0
    if reduce_hist is None and fast and x.numel() >= 65536:
        # single GPU: one full read (sample bracket -> count + collect -> exact select among the candidates).  The
        # status word stays on the device: it arms the 3-pass select, whose kernels return at once when the bracket
        # held (no host read; the whole call is stream-ordered and graph-capturable) - all from ONE library call
        state = torch.empty(h.dlmcq_obs_kth_state_bytes() + 8, dtype=torch.uint8, device=x.device)
        status = state[-8:].view(torch.int32)
        with torch.cuda.device(x.device):
            nws = h.dlmcq_obs_kth_fast_workspace_bytes(x.numel())
            ws = _scratch(x.device, nws)          # NOT the shared zero-contract workspace: this call scribbles on all of it
            _lib.check(h.dlmcq_obs_kth_auto(_ptr(x), x.numel(), _dtype_code(x), flags, ranks[0],
                                            ranks[1] if len(ranks) > 1 else 0, _ptr(values),
                                            C.c_void_p(status.data_ptr()), _ptr(state), _ptr(ws), nws, _stream_ptr()))
        return values[:len(ranks)]
    state = torch.zeros(h.dlmcq_obs_kth_state_bytes(), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        st = _stream_ptr()
        _lib.check(h.dlmcq_obs_kth_begin(_ptr(state), ranks[0], ranks[1] if len(ranks) > 1 else 0, st))
        for p in range(3):
            _lib.check(h.dlmcq_obs_kth_hist(_ptr(x), x.numel(), _dtype_code(x), flags, p, _ptr(state), st))
            if reduce_hist is not None:
                reduce_hist(state[256:].view(torch.int32))
            _lib.check(h.dlmcq_obs_kth_select(p, _ptr(state), st))
        _lib.check(h.dlmcq_obs_kth_values(_ptr(state), _ptr(values), st))
    return values[:len(ranks)]


def sweep_tensor_sse(x, stats, n_bits, allow_offset=True):
    """80 squared-error sums of the per-tensor clip sweep (ops.py:52-61) in one read of x."""
    x = x.detach().contiguous()
    sse = torch.empty(SWEEP_CANDIDATES, dtype=torch.float32, device=x.device)
    lay = layout_of(x)
    with torch.cuda.device(x.device):
        ws, n = _ws_for(x, lay)
        _lib.check(_lib.lib().dlmcq_obs_sweep_tensor_sse(_ptr(x), x.numel(), _dtype_code(x), _ptr(stats), int(n_bits),
                                                         int(bool(allow_offset)), _ptr(sse), _ptr(ws), n, _stream_ptr()))
    return sse


def sweep_tensor_pick(sse, stats, rows_for_mean, n_bits, allow_offset=True):
    scale = torch.empty(1, dtype=torch.float32, device=sse.device)
    offset = torch.empty(1, dtype=torch.float32, device=sse.device)
    picked = torch.empty(1, dtype=torch.int32, device=sse.device)
    with torch.cuda.device(sse.device):
        _lib.check(_lib.lib().dlmcq_obs_sweep_tensor_finalize(_ptr(sse), _ptr(stats), float(rows_for_mean), int(n_bits),
                                                              int(bool(allow_offset)), _ptr(scale), _ptr(offset),
                                                              _ptr(picked), _stream_ptr()))
    return scale, offset, picked


def sweep_tensor(x, n_bits, allow_offset=True, reduce_stats=None, reduce_sse=None):
    """ops.py:36-68 unsigned branch.  reduce_* are optional collectives for multi-GPU callers."""
    _require_cuda(x, "x")
    stats = obs_stats(x)
    if reduce_stats is not None:
        stats = reduce_stats(stats)
    sse = sweep_tensor_sse(x, stats, n_bits, allow_offset)
    rows = x.numel() / x.shape[1] if x.dim() >= 2 else 1.0
    if reduce_sse is not None:
        sse, rows = reduce_sse(sse, rows)
    return sweep_tensor_pick(sse, stats, rows, n_bits, allow_offset)


def sweep_channel(rows2d, n_bits, signed, geom_channels=None):
    """ops.py:169-196 on a contiguous [channels, inner] matrix -> (scale[C], offset[C]).  geom_channels: the
    row count of the whole matrix when rows2d is one rank's block of it (dist.rows_sharded)."""
    _require_cuda(rows2d, "tensor")
    rows2d = rows2d.detach().contiguous()
    ch, inner = rows2d.shape
    scale = torch.empty(ch, dtype=torch.float32, device=rows2d.device)
    offset = torch.empty(ch, dtype=torch.float32, device=rows2d.device)
    with torch.cuda.device(rows2d.device):
        _lib.check(_lib.lib().dlmcq_obs_sweep_channel_geom(_ptr(rows2d), ch, inner, _dtype_code(rows2d), int(n_bits),
                                                           int(bool(signed)), int(geom_channels or ch), _ptr(scale),
                                                           _ptr(offset), _stream_ptr()))
    return scale, offset


def sweep_channel_grouped(rows_list, n_bits, signed):
    """ops.py:169-196 for MANY [channels_i, inner_i] matrices (all weight tensors of a model) in one launch per
    shared-memory class (rows that stage in <= 48 KB per CTA / the rest) instead of one launch per tensor.  Every
    tensor keeps the launch geometry of its own call, so the (scale, offset) vectors are bit-identical to
    sweep_channel(rows_i).  Returns [(scale_i, offset_i)]."""
    if not rows_list:
        return []
    rows_list = [r.detach().contiguous() for r in rows_list]
    dev, dt = rows_list[0].device, rows_list[0].dtype
    for r in rows_list:
        _require_cuda(r, "tensor")
        if r.dim() != 2 or r.device != dev or r.dtype != dt:
            raise DlmcqError("sweep_channel_grouped takes [channels, inner] matrices of one dtype on one device")
    h = _lib.lib()
    total_ch = sum(r.shape[0] for r in rows_list)
    flat = torch.empty(2, total_ch, dtype=torch.float32, device=dev)
    items, outs, c0 = [], [], 0
    for r in rows_list:
        ch, inner = r.shape
        it = _lib.SweepItem()
        it.x, it.channels, it.inner, it.n_bits, it.is_signed = r.data_ptr(), ch, inner, int(n_bits), int(bool(signed))
        it.scale, it.offset = flat[0, c0:].data_ptr(), flat[1, c0:].data_ptr()
        _lib.check(h.dlmcq_obs_sweep_channel_plan(C.byref(it), ch))
        items.append(it)
        outs.append((flat[0, c0:c0 + ch], flat[1, c0:c0 + ch]))
        c0 += ch
    with torch.cuda.device(dev):
        for cls in (lambda b: b <= 48 * 1024, lambda b: b > 48 * 1024):
            sel = [it for it in items if cls(it.smem_bytes)]
            if not sel:
                continue
            arr = (_lib.SweepItem * len(sel))(*sel)
            prefix, acc = [], 0
            for it in sel:
                prefix.append(acc)
                acc += it.ctas
            tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
            pre = torch.tensor(prefix, dtype=torch.int64).to(dev)
            _lib.check(h.dlmcq_obs_sweep_channel_grouped(_ptr(tab), _ptr(pre), len(sel), acc, _dtype_code(rows_list[0]),
                                                         max(it.smem_bytes for it in sel), _stream_ptr()))
    return outs


def l2norm_fixed_point(rows2d, scale, offset, lo, hi, max_iters=1000, poll_every=8, resident=True):
    """ops.py:71-83 / 198-215: iterate s <- sum(x q)/sum(q q + 1e-7) on device until the relative
    change is <= 1e-5.  The loop runs on the device; the host only polls a flag every `poll_every`
    launches.  `max_iters` bounds the search (the reference loops forever on some inputs)."""
    _require_cuda(rows2d, "tensor")
    rows2d = rows2d.detach().contiguous()
    ch, inner = rows2d.shape
    dev = rows2d.device
    scale = _qvec(scale, ch, dev, "scale").clone()
    offset = _qvec(offset, ch, dev, "offset")
    diff = torch.zeros(1, dtype=torch.float32, device=dev)
    flags = torch.zeros(2, dtype=torch.int32, device=dev)       # [done, iters]
    lay = Layout(1, ch, inner, _dtype_code(rows2d))
    h = _lib.lib()
    with torch.cuda.device(dev):
        ws, n = _ws_for(rows2d, lay)
        if resident:
            # the whole loop in one cooperative launch when the tensor fits in shared memory (every CNN weight matrix)
            st = h.dlmcq_obs_l2norm_resident(_ptr(rows2d), ch, inner, lay.dtype, _ptr(scale), _ptr(offset), int(lo),
                                             int(hi), int(max_iters), _ptr(diff), C.c_void_p(flags.data_ptr()),
                                             C.c_void_p(flags.data_ptr() + 4), _ptr(ws), n, _stream_ptr())
            if st == 0:
                f = flags.tolist()
                return scale, int(f[1]), bool(f[0])
            if st != -5:                        # anything but "not resident": a real error
                _lib.check(st)
        it = 0
        while it < max_iters:
            for _ in range(min(poll_every, max_iters - it)):
                _lib.check(h.dlmcq_obs_l2norm_step(_ptr(rows2d), ch, inner, lay.dtype, _ptr(scale), _ptr(offset),
                                                   int(lo), int(hi), _ptr(diff), C.c_void_p(flags.data_ptr()),
                                                   C.c_void_p(flags.data_ptr() + 4), _ptr(ws), n, _stream_ptr()))
                it += 1
            if int(flags[0].item()):
                break
    return scale, int(flags[1].item()), bool(flags[0].item())


# --------------------------------------------------------------------------------------
# AdaRound
def adaround_forward(w, alpha, scale, lo, hi, soft, ch_axis=0):
    _require_cuda(w, "w")
    w, alpha = w.detach().contiguous(), alpha.detach().contiguous()
    lay = layout_of(w, ch_axis)
    s = _qvec(scale, lay.channels, w.device, "scale")
    y = torch.empty_like(w)
    with torch.cuda.device(w.device):
        _lib.check(_lib.lib().dlmcq_adaround_forward(_ptr(w), _ptr(alpha), _ptr(y), C.byref(lay), _ptr(s), int(lo),
                                                     int(hi), int(bool(soft)), _stream_ptr()))
    return y


def adaround_backward(w, alpha, dy, scale, lo, hi, ch_axis=0):
    w, alpha, dy = w.detach().contiguous(), alpha.detach().contiguous(), dy.detach().contiguous()
    lay = layout_of(w, ch_axis)
    s = _qvec(scale, lay.channels, w.device, "scale")
    dalpha = torch.empty_like(w)
    ds = torch.empty(lay.channels, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        ws, n = _ws_for(w, lay)
        _lib.check(_lib.lib().dlmcq_adaround_backward(_ptr(w), _ptr(alpha), _ptr(dy), _ptr(dalpha), _ptr(ds),
                                                      C.byref(lay), _ptr(s), int(lo), int(hi), _ptr(ws), n,
                                                      _stream_ptr()))
    return dalpha, ds


def adaround_init_alpha(w, scale, ch_axis=0):
    _require_cuda(w, "w")
    w = w.detach().contiguous()
    lay = layout_of(w, ch_axis)
    s = _qvec(scale, lay.channels, w.device, "scale")
    alpha = torch.empty_like(w)
    with torch.cuda.device(w.device):
        _lib.check(_lib.lib().dlmcq_adaround_init_alpha(_ptr(w), _ptr(alpha), C.byref(lay), _ptr(s), _stream_ptr()))
    return alpha


# --------------------------------------------------------------------------------------
# RootQ
def _f32_scalar(t, device, name):
    if t.device != device or t.dtype != torch.float32:
        raise DlmcqError(f"{name} must be a float32 tensor on {device}")
    return t


def rootq_act_prepare(in_scale, run_scale, momentum, g, lo, hi, training):
    """RootQ/base.py:92-106.  Updates run_scale in place when training; returns the state block."""
    dev = run_scale.device
    state = torch.empty(ROOTQ_STATE_FLOATS, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dlmcq_rootq_act_prepare(_ptr(_f32_scalar(in_scale.detach(), dev, "in_scale")),
                                                      _ptr(_f32_scalar(run_scale, dev, "run_scale")), float(momentum),
                                                      float(g), int(lo), int(hi), int(bool(training)), _ptr(state),
                                                      _stream_ptr()))
    return state


def rootq_wt_prepare(upper, lower, alpha, run_upper, run_lower, momentum, g, lo, hi, training):
    """RootQ/base.py:131-147 (+ alpha clamp of function.py:25-26)."""
    dev = run_upper.device
    state = torch.empty(ROOTQ_STATE_FLOATS, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dlmcq_rootq_wt_prepare(_ptr(_f32_scalar(upper.detach(), dev, "wt_upper")),
                                                     _ptr(_f32_scalar(lower.detach(), dev, "wt_lower")),
                                                     _ptr(_f32_scalar(alpha.detach(), dev, "wt_alpha")),
                                                     _ptr(run_upper), _ptr(run_lower), float(momentum), float(g),
                                                     int(lo), int(hi), int(bool(training)), _ptr(state), _stream_ptr()))
    return state


def _rootq_fwd(fn_name, x, state):
    _require_cuda(x, "x")
    x = x.detach().contiguous()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(getattr(_lib.lib(), fn_name)(_ptr(x), _ptr(y), x.numel(), _dtype_code(x), _ptr(state), _stream_ptr()))
    return y


def _rootq_bwd(fn_name, x, dy, state, ngrads):
    x, dy = x.detach().contiguous(), dy.detach().contiguous()
    dx = torch.empty_like(x)
    grads = torch.empty(ngrads, dtype=torch.float32, device=x.device)
    lay = layout_of(x)
    with torch.cuda.device(x.device):
        ws, n = _ws_for(x, lay)
        _lib.check(getattr(_lib.lib(), fn_name)(_ptr(x), _ptr(dy), _ptr(dx), _ptr(grads), x.numel(), _dtype_code(x),
                                                _ptr(state), _ptr(ws), n, _stream_ptr()))
    return dx, grads


def rootq_act_forward(x, state):
    return _rootq_fwd("dlmcq_rootq_act_forward", x, state)


def rootq_wt_forward(w, state):
    return _rootq_fwd("dlmcq_rootq_wt_forward", w, state)


def rootq_act_backward(x, dy, state):
    """-> dx, d_in_scale[1]"""
    return _rootq_bwd("dlmcq_rootq_act_backward", x, dy, state, 1)


def rootq_wt_backward(w, dy, state):
    """-> dw, grads[3] = (d wt_upper, d wt_lower, d wt_alpha)"""
    return _rootq_bwd("dlmcq_rootq_wt_backward", w, dy, state, 3)


class DeferredScaleGrads:
    """Per-tensor backward with the scale-gradient reduction deferred to ONE launch for many tensors
    (dlmcq_fq_backward_partials + dlmcq_fq_finalize_many): saves the ~3.4 us serial in-kernel finalisation
    of every layer; results are identical (same fixed-order double summation of the same partials)."""

    def __init__(self, device, capacity):
        self.device = torch.device(device)
        self.floats = _lib.lib().dlmcq_fq_partials_floats()
        self.partials = torch.empty(capacity, self.floats, dtype=torch.float32, device=self.device)
        self.capacity, self._table, self._key = capacity, None, None

    def backward(self, slot, x, dy, dx, scale, offset, lo, hi, form, g=0.0):
        """Enqueue the backward of one contiguous tensor; its partials go to `slot`."""
        lay = layout_of(x)
        qp = QParams(form, int(lo), int(hi), float(g), scale.data_ptr(), offset.data_ptr() if offset is not None else None)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_fq_backward_partials(_ptr(x), _ptr(dy), _ptr(dx), C.byref(lay), C.byref(qp),
                                                             _ptr(self.partials[slot]), _stream_ptr()))

    def finalize(self, dscales):
        """dscales: list of 1-element float32 device tensors, one per slot 0..len-1."""
        key = tuple(d.data_ptr() for d in dscales)
        if key != self._key:
            arr = (_lib.FinalizeItem * len(dscales))()
            for i, d in enumerate(dscales):
                arr[i].partials, arr[i].dscale = self.partials[i].data_ptr(), d.data_ptr()
            self._table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
            self._key = key
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_fq_finalize_many(_ptr(self._table), len(dscales), _stream_ptr()))


# --------------------------------------------------------------------------------------
GROUP_SEG = 4096


class GroupedFakeQuant:
    """All weight tensors of a model in one launch (forward) / two launches (backward).

    The descriptor table lives on the device and is rebuilt only when the set of tensors changes,
    so a training step costs one small H2D copy at most (none when pointers are stable)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._key = None
        self._items = self._unit_prefix = self._chan_prefix = self._partials = None
        self._total_units = self._total_channels = 0

    def _table(self, entries, backward):
        key = tuple((e["x"].data_ptr(), e["y"].data_ptr(), e["dy"].data_ptr() if backward else 0, e["scale"].data_ptr(),
                     e["offset"].data_ptr() if e.get("offset") is not None else 0,
                     e["dscale"].data_ptr() if backward else 0, e["channels"], e["inner"], e["form"], e["lo"], e["hi"],
                     e["g"]) for e in entries)
        if key == self._key:
            return
        arr = (_lib.GroupItem * len(entries))()
        units, chans = [0], [0]
        for i, e in enumerate(entries):
            it = arr[i]
            it.x, it.y = e["x"].data_ptr(), e["y"].data_ptr()
            it.dy = e["dy"].data_ptr() if backward else None
            it.scale = e["scale"].data_ptr()
            it.offset = e["offset"].data_ptr() if e.get("offset") is not None else None
            it.dscale = e["dscale"].data_ptr() if backward else None
            it.channels, it.inner = e["channels"], e["inner"]
            it.form, it.lo, it.hi, it.g = e["form"], e["lo"], e["hi"], e["g"]
            units.append(units[-1] + e["channels"] * ((e["inner"] + GROUP_SEG - 1) // GROUP_SEG))
            chans.append(chans[-1] + e["channels"])
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._items = raw.to(self.device)
        self._unit_prefix = torch.tensor(units, dtype=torch.int64).to(self.device)
        self._chan_prefix = torch.tensor(chans, dtype=torch.int64).to(self.device)
        self._total_units, self._total_channels = units[-1], chans[-1]
        if self._partials is None or self._partials.numel() < self._total_units:
            self._partials = torch.empty(max(self._total_units, 1), dtype=torch.float32, device=self.device)
        self._key = key

    def forward(self, entries, dtype=torch.float32):
        """entries: dicts with x, y, scale, offset|None, channels, inner, form, lo, hi, g."""
        self._table(entries, backward=False)
        code = F32 if dtype == torch.float32 else BF16
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_fq_forward_grouped(_ptr(self._items), _ptr(self._unit_prefix), len(entries),
                                                           self._total_units, code, _stream_ptr()))

    def backward(self, entries, dtype=torch.float32):
        """entries additionally carry dy, dscale; `y` receives dx."""
        self._table(entries, backward=True)
        code = F32 if dtype == torch.float32 else BF16
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_fq_backward_grouped(_ptr(self._items), _ptr(self._unit_prefix),
                                                            _ptr(self._chan_prefix), len(entries), self._total_units,
                                                            self._total_channels, code, _ptr(self._partials),
                                                            _stream_ptr()))


class GroupedRootQ:
    """RootQ for a whole model with a constant number of launches: ONE launch prepares every quantizer's scalar
    state (EMA of the running bounds, gradient mix, delta, clamped alpha - RootQ/base.py:92-101,131-147, running
    buffers updated in place), ONE quantises all weight tensors, ONE (+ a finalisation) differentiates them.
    Descriptor tables live on the device and are rebuilt only when pointers change."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._prep_key = self._prep_tab = self.states = None
        self._wt = {(w, b): [None, None, None, 0] for w in ("wt", "act") for b in (False, True)}   # key, items, prefix, units
        self._partials = {"wt": None, "act": None}

    def prepare(self, quantizers):
        """quantizers: dicts with kind 'act' (in_scale, run_scale) or 'wt' (upper, lower, alpha, run_upper,
        run_lower), plus momentum, g, lo, hi, training.  Returns the list of 8-float state views."""
        key = tuple((q["kind"], *(q[k].data_ptr() for k in (("in_scale", "run_scale") if q["kind"] == "act" else
                                                             ("upper", "lower", "alpha", "run_upper", "run_lower"))),
                     q["momentum"], q["g"], q["lo"], q["hi"], bool(q["training"])) for q in quantizers)
        if key != self._prep_key:
            n = len(quantizers)
            self.states = torch.zeros(n, ROOTQ_STATE_FLOATS, dtype=torch.float32, device=self.device)
            arr = (_lib.RootqPrep * n)()
            for i, q in enumerate(quantizers):
                it = arr[i]
                if q["kind"] == "act":
                    it.param_a, it.run_a, it.is_weight = q["in_scale"].data_ptr(), q["run_scale"].data_ptr(), 0
                else:
                    it.param_a, it.param_b, it.alpha = q["upper"].data_ptr(), q["lower"].data_ptr(), q["alpha"].data_ptr()
                    it.run_a, it.run_b, it.is_weight = q["run_upper"].data_ptr(), q["run_lower"].data_ptr(), 1
                it.state = self.states[i].data_ptr()
                it.momentum, it.g, it.lo, it.hi = float(q["momentum"]), float(q["g"]), int(q["lo"]), int(q["hi"])
                it.training = int(bool(q["training"]))
            self._prep_tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
            self._prep_key = key
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_rootq_prepare_many(_ptr(self._prep_tab), len(quantizers), _stream_ptr()))
        return [self.states[i] for i in range(len(quantizers))]

    def _table(self, entries, backward, kind="wt"):
        slot = self._wt[(kind, backward)]
        key = tuple((e["w"].data_ptr(), e["out"].data_ptr(), e["dy"].data_ptr() if backward else 0,
                     e["state"].data_ptr(), e["grads"].data_ptr() if backward else 0, e["w"].numel()) for e in entries)
        if key == slot[0]:
            return slot
        arr = (_lib.RootqItem * len(entries))()
        units = [0]
        for i, e in enumerate(entries):
            it = arr[i]
            it.x, it.y, it.state, it.numel = e["w"].data_ptr(), e["out"].data_ptr(), e["state"].data_ptr(), e["w"].numel()
            it.dy = e["dy"].data_ptr() if backward else None
            it.grads = e["grads"].data_ptr() if backward else None
            units.append(units[-1] + (e["w"].numel() + _lib.ROOTQ_UNIT - 1) // _lib.ROOTQ_UNIT)
        slot[0] = key
        slot[1] = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        slot[2] = torch.tensor(units, dtype=torch.int64).to(self.device)
        slot[3] = units[-1]
        if self._partials[kind] is None or self._partials[kind].numel() < 3 * units[-1]:
            self._partials[kind] = torch.empty(max(3 * units[-1], 1), dtype=torch.float32, device=self.device)
        return slot

    def wt_forward(self, entries, dtype=torch.float32):
        """entries: dicts with w, out (receives w_q), state (contiguous tensors)."""
        _, items, prefix, units = self._table(entries, False)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_rootq_wt_forward_grouped(_ptr(items), _ptr(prefix), len(entries), units,
                                                                 F32 if dtype == torch.float32 else BF16, _stream_ptr()))

    def wt_backward(self, entries, dtype=torch.float32):
        """entries additionally carry dy and grads [3] (d wt_upper, d wt_lower, d wt_alpha); out receives dw."""
        _, items, prefix, units = self._table(entries, True)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_rootq_wt_backward_grouped(_ptr(items), _ptr(prefix), len(entries), units,
                                                                  F32 if dtype == torch.float32 else BF16,
                                                                  _ptr(self._partials["wt"]), _stream_ptr()))

    def act_forward(self, entries, dtype=torch.float32):
        """Several ACTIVATION tensors in one launch.  entries: dicts with w (the activation), out (receives x_q), state
        (the block `prepare` / rootq_act_prepare wrote) - contiguous tensors resident at the same time: a block input
        that feeds two quantised convolutions, cached calibration activations, a quantizer-set measurement."""
        _, items, prefix, units = self._table(entries, False, "act")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_rootq_act_forward_grouped(_ptr(items), _ptr(prefix), len(entries), units,
                                                                  F32 if dtype == torch.float32 else BF16, _stream_ptr()))

    def act_backward(self, entries, dtype=torch.float32):
        """entries additionally carry dy and grads [1] (d in_scale); out receives dx."""
        _, items, prefix, units = self._table(entries, True, "act")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_rootq_act_backward_grouped(_ptr(items), _ptr(prefix), len(entries), units,
                                                                   F32 if dtype == torch.float32 else BF16,
                                                                   _ptr(self._partials["act"]), _stream_ptr()))


class HostFakeQuant:
    """End-to-end path for HOST tensors (pinned memory) on a caller-owned `dlmcq_host_ctx`: forward + backward with the
    H2D / D2H copies pipelined against the kernels.  All calls enqueue; `synchronize()` waits.

    forward_backward*  y, dx come back in the tensor's dtype;
    codes_async        compact lossless result: packed integer codes (y = code * s' + offset) and one keep bit per
                       element (dx = keep ? dy : 0) - 8.6 instead of 16 bytes per fp32 element over PCIe."""

    def __init__(self, device, chunk_elems=1 << 22, dtype=torch.float32):
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.chunk = int(chunk_elems)
        self.code = F32 if dtype == torch.float32 else BF16
        self._ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dlmcq_host_ctx_create(C.byref(self._ctx), self.chunk))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            _lib.lib().dlmcq_host_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _host(*tensors):
        for t in tensors:
            if t is not None and (t.is_cuda or not t.is_contiguous()):
                raise DlmcqError("host path takes contiguous CPU tensors")

    def forward_backward_async(self, x, dy, y, dx, dscale_out, scale, offset, lo, hi, form=FORM_AFFINE, g=0.0):
        """Enqueue only; `dscale_out` is a 1-element pinned float32 tensor that receives the scale gradient."""
        self._host(x, dy, y, dx, dscale_out)
        _lib.check(_lib.lib().dlmcq_host_ctx_fq_forward_backward(
            self._ctx, _ptr(x), _ptr(dy), _ptr(y), _ptr(dx), C.c_void_p(dscale_out.data_ptr()), x.numel(), self.code,
            int(form), int(lo), int(hi), float(g), float(scale), float(offset)))

    def forward_backward(self, x, dy, y, dx, scale, offset, lo, hi, form=FORM_AFFINE, g=0.0):
        """x, dy: host inputs; y, dx: host outputs (same shape/dtype).  Returns dscale (python float)."""
        ds = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.forward_backward_async(x, dy, y, dx, ds, scale, offset, lo, hi, form, g)
        self.synchronize()
        return float(ds)

    def codes_async(self, x, dy, codes, keep, dscale_out, scale, offset, lo, hi, form=FORM_AFFINE, g=0.0, pack4=False):
        """codes: uint8/int8 host tensor of numel (or ceil(numel/2) with pack4) bytes; keep: uint8 host tensor of
        ceil(numel/8) bytes (dy, keep, dscale_out may be None: forward only)."""
        self._host(x, dy, codes, keep, dscale_out)
        _lib.check(_lib.lib().dlmcq_host_ctx_fq_codes(
            self._ctx, _ptr(x), _ptr(dy), _ptr(codes), _ptr(keep),
            C.c_void_p(dscale_out.data_ptr()) if dscale_out is not None else None, x.numel(), self.code, int(form),
            int(lo), int(hi), float(g), float(scale), float(offset), int(bool(pack4))))

    def synchronize(self):
        _lib.check(_lib.lib().dlmcq_host_ctx_synchronize(self._ctx))
