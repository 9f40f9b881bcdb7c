"""The layer's matrix product on the integer codes (consumer side).

Reference: `_forward_func(q_input, q_weight)` - F.linear / F.conv2d on the two FAKE-QUANTISED fp32 tensors
(dlmc/quantization/scalar/modules/linear.py, modules/conv.py, entered from modules/base.py:140 and
FSPTQuant/base.py:111-113).  For a per-tensor activation quantizer and a per-output-channel (or per-tensor) weight
quantizer without an additive weight term that product factors into an exact integer dot product of the codes and a
per-output-channel affine map (include/dlmcq.h, `dlmcq_qgemm`).  The integer product runs on the tcgen05 tensor cores
(csrc/qgemm_kernels.cu); the activation is read at one byte per element instead of four.

Two levels:
  * functional: `codes_forward`, `qgemm_prepare`, `qgemm` - thin wrappers over the C ABI;
  * module: `enable_code_gemm(model)` switches every eligible quantised `nn.Linear` / 1x1 `nn.Conv2d` of a model to
    the code path for inference (`module.eval()`, no autograd); training / calibration / ineligible layers keep the
    reference's path (fake-quant kernels + the library convolution).  Opt-in: results equal the fake-quant path up
    to the rounding of the fp32 product (the integer accumulation itself is exact; cuDNN's TF32 / fp32 sums are not).
"""
import ctypes as C
import math
import types

import torch
from torch import nn

from . import _lib
from . import functional as F
from ._lib import BF16, F32, FORM_AFFINE, FORM_SYM, FORM_ZP, QGEMM_E4M3, QGEMM_I8, DlmcqError, QParams

__all__ = ["codes_forward", "qgemm_prepare", "qgemm", "enable_code_gemm", "disable_code_gemm", "QGEMM_I8", "QGEMM_E4M3"]


def _qp(form, lo, hi, g, scale, offset):
    return QParams(int(form), int(lo), int(hi), float(g), scale.data_ptr(), offset.data_ptr() if offset is not None else None)


def codes_forward(x, scale, offset, lo, hi, form, g=0.0, ch_axis=None, encoding=QGEMM_I8):
    """x -> its integer codes, one byte per element (uint8 tensor of x's shape and memory format; two's complement
    for signed ranges, or the e4m3 byte of the code with `encoding=QGEMM_E4M3`).  Same codes as `fq_forward`."""
    F._require_cuda(x, "x")
    x = F._dense(x.detach(), ch_axis)
    lay = F.layout_of(x, ch_axis)
    s = F._qvec(scale, lay.channels, x.device, "scale")
    o = F._qvec(offset, lay.channels, x.device, "offset")
    codes = torch.empty_like(x, dtype=torch.uint8)
    qp = _qp(form, lo, hi, g, s, o)
    with F._on(x.device):
        _lib.check(_lib.lib().dlmcq_codes_forward(F._ptr(x), F._ptr(codes), C.byref(lay), C.byref(qp), int(encoding),
                                                  F._stream_ptr()))
    return codes


def qgemm_prepare(w_codes, act, wt, bias=None, encoding=QGEMM_I8):
    """alpha[n], beta[n] of `qgemm` from device-resident quantizer parameters (no host sync).
    act = (scale, offset_or_zero_point, lo, hi, form, g), wt = (scale [n] or [1], lo, hi, form, g); w_codes [n, k]."""
    F._require_cuda(w_codes, "w_codes")
    if w_codes.dtype != torch.uint8 or w_codes.dim() != 2 or not w_codes.is_contiguous():
        raise DlmcqError("w_codes must be a contiguous [n, k] uint8 tensor")
    n, k = w_codes.shape
    dev = w_codes.device
    a_s, a_o, a_lo, a_hi, a_form, a_g = act
    w_s, w_lo, w_hi, w_form, w_g = wt
    a_s = F._qvec(a_s, 1, dev, "activation scale")
    a_o = F._qvec(a_o, 1, dev, "activation offset")
    w_ch = n if (isinstance(w_s, torch.Tensor) and w_s.numel() == n and n != 1) else 1
    w_s = F._qvec(w_s, w_ch, dev, "weight scale")
    b = None
    if bias is not None:
        b = bias.detach().to(device=dev, dtype=torch.float32).contiguous()
        if b.numel() != n:
            raise DlmcqError("bias must have n entries")
    alpha = torch.empty(n, dtype=torch.float32, device=dev)
    beta = torch.empty(n, dtype=torch.float32, device=dev)
    qa, qw = _qp(a_form, a_lo, a_hi, a_g, a_s, a_o), _qp(w_form, w_lo, w_hi, w_g, w_s, None)
    with F._on(dev):
        _lib.check(_lib.lib().dlmcq_qgemm_prepare(F._ptr(w_codes), n, k, int(encoding), C.byref(qa), C.byref(qw), w_ch,
                                                  F._ptr(b), F._ptr(alpha), F._ptr(beta), F._stream_ptr()))
    return alpha, beta


def qgemm(a_codes, w_codes, alpha, beta, relu=False, out_dtype=torch.float32, a_signed=False, encoding=QGEMM_I8,
          out=None):
    """out[m, n] = alpha[n] * (sum_k a_codes[m, k] * w_codes[n, k]) + beta[n]  (two fp32 roundings), optional ReLU."""
    F._require_cuda(a_codes, "a_codes")
    if a_codes.dtype != torch.uint8 or w_codes.dtype != torch.uint8 or a_codes.dim() != 2 or w_codes.dim() != 2:
        raise DlmcqError("a_codes [m, k] and w_codes [n, k] must be uint8 matrices")
    if not (a_codes.is_contiguous() and w_codes.is_contiguous()):
        raise DlmcqError("code matrices must be contiguous (row-major, K innermost)")
    m, k = a_codes.shape
    n, k2 = w_codes.shape
    if k != k2 or alpha.numel() != n or beta.numel() != n:
        raise DlmcqError("shape mismatch between a_codes, w_codes, alpha, beta")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise DlmcqError("out_dtype must be float32 or bfloat16")
    if out is None:
        out = torch.empty(m, n, dtype=out_dtype, device=a_codes.device)
    elif out.shape != (m, n) or out.dtype != out_dtype or not out.is_contiguous() or out.device != a_codes.device:
        raise DlmcqError("out must be a contiguous [m, n] tensor of out_dtype on the operands' device")
    with F._on(a_codes.device):
        _lib.check(_lib.lib().dlmcq_qgemm(F._ptr(a_codes), F._ptr(w_codes), F._ptr(alpha), F._ptr(beta), F._ptr(out), m, n,
                                          k, int(encoding), int(bool(a_signed)), int(bool(relu)),
                                          F32 if out_dtype == torch.float32 else BF16, F._stream_ptr()))
    return out


# ---------------------------------------------------------------------------------------------------------------
# module level
# ---------------------------------------------------------------------------------------------------------------
def _family(mod):
    """(activation form, weight form) of a quantised layer, or None for families the code path does not serve."""
    from .scalar.FSPTQuant.base import FSPTQBase
    from .scalar.modules.base import QBase
    if isinstance(mod, QBase):
        return FORM_AFFINE, FORM_AFFINE
    if isinstance(mod, FSPTQBase):
        if mod.qconfig['weight'].get('recon_type') in ('adaround', 'dist_recon'):
            return None                      # AdaRound codes come from a different rounding rule (adaround_kernels.cu)
        return FORM_ZP, FORM_SYM
    return None


def _geometry_ok(mod):
    if isinstance(mod, nn.Linear):
        return mod.in_features % 16 == 0
    if isinstance(mod, nn.Conv2d):
        one = (1, 1)
        return (tuple(mod.kernel_size) == one and tuple(mod.stride) == one and tuple(mod.dilation) == one and
                tuple(mod.padding) == (0, 0) and mod.groups == 1 and mod.padding_mode == 'zeros' and
                mod.in_channels % 16 == 0)
    return False


def _enabled(mod):
    q = mod.qconfig
    if hasattr(mod, 'act_quant'):            # FSPTQBase switches (change_quant_state)
        return bool(mod.act_quant) and bool(mod.wt_quant)
    return bool(q['input']['enable']) and bool(q['weight']['enable'])


class _CodeGemmState:
    """Per-layer cache of the weight codes and the epilogue vectors, rebuilt when any tensor they derive from changes."""

    def __init__(self, encoding):
        self.encoding = encoding
        self.key = None
        self.w_codes = self.alpha = self.beta = None
        self.usable = None                   # None = not decided yet (needs initialised qparams)

    def refresh(self, mod, forms):
        a_form, w_form = forms
        off = getattr(mod, 'wt_offset', None)
        key = (mod.weight._version, mod.wt_scale._version, mod.in_scale._version,
               None if mod.in_offset is None else (mod.in_offset.data_ptr(), mod.in_offset._version),
               None if mod.bias is None else mod.bias._version, mod.weight.data_ptr())
        if key == self.key:
            return self.usable
        self.key = key
        n = mod.weight.shape[0]
        ws, ins = mod.wt_scale.detach(), mod.in_scale.detach()
        usable = ins.numel() == 1 and ws.numel() in (1, n)
        if usable and w_form == FORM_AFFINE and isinstance(off, torch.Tensor):
            usable = not bool(off.any())     # one sync per weight update, never per forward
        if usable and a_form == FORM_ZP and isinstance(mod.in_offset, torch.Tensor):
            # FSPTQuant/base.py:108-109 adds whatever offset the observer returned; only an INTEGRAL zero-point keeps
            # clamp(round(x/s) + zp) an integer code (the min/max observer of a post-ReLU tensor gives 0)
            zp = mod.in_offset.detach()
            usable = bool((zp == zp.round()).all())
        lo_w, hi_w, lo_a, hi_a = mod.wt_min_val, mod.wt_max_val, mod.in_min_val, mod.in_max_val
        if self.encoding == QGEMM_E4M3:
            usable = usable and lo_w >= -16 and hi_w <= 16 and lo_a >= -16 and hi_a <= 16
        else:
            usable = usable and lo_w >= -128 and hi_w <= 127 and lo_a >= -128 and hi_a <= 255 and (lo_a >= 0 or hi_a <= 127)
        self.usable = usable
        if not usable:
            self.w_codes = self.alpha = self.beta = None
            return False
        w = mod.weight.detach()
        w2 = w.reshape(n, -1)                # 1x1 conv weights [n, k, 1, 1] are [n, k] matrices in either memory format
        w2 = w2 if w2.is_contiguous() else w2.contiguous()
        g_w = 1 / math.sqrt(w.numel() * hi_w) if w_form == FORM_AFFINE else 0.0
        per_ch = ws.numel() == n and n != 1
        self.w_codes = codes_forward(w2, ws.reshape(-1), None, lo_w, hi_w, w_form, g_w, ch_axis=0 if per_ch else None,
                                     encoding=self.encoding)
        self._wt = (ws.reshape(-1), lo_w, hi_w, w_form, g_w)
        self._g_a = 0.0                      # the activation's grad_scale factor depends on the input's numel: epilogue()
        self.alpha = self.beta = None
        self._ab_numel = None
        return True

    def epilogue(self, mod, forms, numel):
        """alpha / beta; for the AFFINE activation form the multiplier is the grad_scale value (s - s*g) + s*g with
        g = 1/sqrt(numel * qmax) (modules/base.py:96-97), so the vectors are keyed on the input's element count."""
        a_form, _ = forms
        if self.alpha is not None and (a_form != FORM_AFFINE or self._ab_numel == numel):
            return self.alpha, self.beta
        g_a = 1 / math.sqrt(numel * mod.in_max_val) if a_form == FORM_AFFINE else 0.0
        act = (mod.in_scale.detach().reshape(-1), mod.in_offset, mod.in_min_val, mod.in_max_val, a_form, g_a)
        self.alpha, self.beta = qgemm_prepare(self.w_codes, act, self._wt, mod.bias, self.encoding)
        self._ab_numel = numel
        self._g_a = g_a
        return self.alpha, self.beta


def _code_forward(self, input):
    st = self.__dict__.get('_code_gemm')
    forms = self.__dict__.get('_code_gemm_forms')
    # inference only: under autograd the layer keeps the differentiable fake-quant path
    if (st is None or self.training or torch.is_grad_enabled() or not input.is_cuda
            or input.dtype not in (torch.float32, torch.bfloat16) or not _enabled(self)
            or not (self._ready('in', self.in_init_state) and self._ready('wt', self.wt_init_state))):
        return self._code_gemm_orig_forward(input)
    if getattr(input, '_dlmcq_q', None) is not None:
        return self._code_gemm_orig_forward(input)      # a fused producer already handed over the fake-quantised tensor
    if not st.refresh(self, forms):
        return self._code_gemm_orig_forward(input)
    a_form, _ = forms
    self.__dict__.pop('_wq', None)                       # a weight group's pre-hook may have staged its copy: unused here
    if isinstance(self, nn.Conv2d):
        if input.dim() != 4:
            return self._code_gemm_orig_forward(input)
        b, c, h, w = input.shape
        x = input if input.is_contiguous(memory_format=torch.channels_last) else \
            input.contiguous(memory_format=torch.channels_last)
        rows = b * h * w
    else:
        x = input if input.is_contiguous() else input.contiguous()
        rows = x.numel() // x.shape[-1]
    numel = x.numel()
    alpha, beta = st.epilogue(self, forms, numel)
    a_codes = codes_forward(x, self.in_scale.detach().reshape(-1), self.in_offset, self.in_min_val, self.in_max_val,
                            a_form, st._g_a, encoding=st.encoding)
    k = self.weight.shape[1] if isinstance(self, nn.Conv2d) else self.weight.shape[-1]
    # elementwise codes keep x's memory order: channels-last NCHW storage is the [B*H*W, C] matrix
    a2 = a_codes.permute(0, 2, 3, 1).reshape(rows, k) if isinstance(self, nn.Conv2d) else a_codes.reshape(rows, k)
    out = qgemm(a2, st.w_codes, alpha, beta, out_dtype=input.dtype, a_signed=self.in_min_val < 0, encoding=st.encoding)
    n = self.weight.shape[0]
    if isinstance(self, nn.Conv2d):
        return out.view(b, h, w, n).permute(0, 3, 1, 2)          # NCHW view of channels-last storage
    return out.view(*input.shape[:-1], n)


def _auto_encoding(mod):
    """e4m3 operands (fp32 accumulators: no int -> float conversion in the epilogue) when every code of the layer is
    exactly representable (|code| <= 16, i.e. up to 4-bit unsigned / 5-bit signed ranges), the integer kind otherwise.
    Both are bit-identical where both apply (tests/qgemm_cases.py)."""
    small = mod.wt_min_val >= -16 and mod.wt_max_val <= 16 and mod.in_min_val >= -16 and mod.in_max_val <= 16
    return QGEMM_E4M3 if small else QGEMM_I8


def enable_code_gemm(model, encoding=None):
    """Route every eligible quantised Linear / 1x1 Conv2d of `model` through the integer-code GEMM in eval mode
    (`encoding`: QGEMM_I8, QGEMM_E4M3, or None = per layer by code range).
    Returns the list of module names switched.  Eligibility that depends on calibrated state (per-tensor activation
    scale, zero weight offset, ...) is re-checked lazily and a layer that fails it silently keeps its original path."""
    names = []
    for name, mod in model.named_modules():
        forms = _family(mod)
        if forms is None or not _geometry_ok(mod) or '_code_gemm' in mod.__dict__:
            continue
        mod.__dict__['_code_gemm'] = _CodeGemmState(_auto_encoding(mod) if encoding is None else encoding)
        mod.__dict__['_code_gemm_forms'] = forms
        mod.__dict__['_code_gemm_orig_forward'] = mod.forward
        mod.forward = types.MethodType(_code_forward, mod)
        names.append(name)
    return names


def disable_code_gemm(model):
    for mod in model.modules():
        if '_code_gemm' in mod.__dict__:
            del mod.__dict__['forward']
            for k in ('_code_gemm', '_code_gemm_forms', '_code_gemm_orig_forward'):
                mod.__dict__.pop(k, None)
