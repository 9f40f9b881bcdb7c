"""TEST INFRASTRUCTURE ONLY - ctypes loader for the plain-C oracle (oracle/fq_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


def build():
    src = os.path.join(HERE, "fq_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "_build/liboracle.so"], check=True, capture_output=True)
    return LIB


_h = None


def lib():
    global _h
    if _h is None:
        _h = C.CDLL(build())
        _h.orc_grad_scale_value.restype = C.c_float
        _h.orc_grad_scale_value.argtypes = [C.c_float, C.c_float]
        _h.orc_rootq_mix.restype = C.c_float
        _h.orc_rootq_mix.argtypes = [C.c_float, C.c_float, C.c_double, C.c_double]
        _h.orc_sweep_tensor.restype = C.c_int
        _h.orc_l2norm_tensor.restype = C.c_int
    return _h


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def fq_forward(x, scale, offset, form, lo, hi, g=0.0, channels=1, inner=None):
    x = _f(x)
    n = x.size
    inner = n if inner is None else inner
    s, o = _f(scale).reshape(-1), (None if offset is None else _f(offset).reshape(-1))
    y, codes = np.empty_like(x), np.empty_like(x)
    lib().orc_fq_forward(_p(x), _p(y), _p(codes), C.c_int64(n), C.c_int64(channels), C.c_int64(inner), _p(s), _p(o),
                         C.c_int(form), C.c_float(lo), C.c_float(hi), C.c_float(g))
    return y, codes


def fq_backward(x, dy, scale, offset, form, lo, hi, g=0.0, channels=1, inner=None):
    x, dy = _f(x), _f(dy)
    n = x.size
    inner = n if inner is None else inner
    s, o = _f(scale).reshape(-1), (None if offset is None else _f(offset).reshape(-1))
    dx, ds = np.empty_like(x), np.zeros(channels, dtype=np.float64)
    lib().orc_fq_backward(_p(x), _p(dy), _p(dx), _p(ds), C.c_int64(n), C.c_int64(channels), C.c_int64(inner), _p(s),
                          _p(o), C.c_int(form), C.c_float(lo), C.c_float(hi), C.c_float(g))
    return dx, ds


def rootq_act(x, dy, rs, up, q):
    x = _f(x)
    dyv = None if dy is None else _f(dy)
    y, dx, d = np.empty_like(x), np.empty_like(x), C.c_double(0.0)
    lib().orc_rootq_act(_p(x), _p(dyv), _p(y), _p(dx), C.byref(d), C.c_int64(x.size), C.c_float(rs), C.c_float(up),
                        C.c_float(q))
    return y, dx, d.value


def rootq_mix(run, param, momentum, g):
    return lib().orc_rootq_mix(C.c_float(run), C.c_float(param), C.c_double(momentum), C.c_double(g))


def rootq_wt(w, dy, U, L, alpha, q):
    w = _f(w)
    dyv = None if dy is None else _f(dy)
    y, dw, gr = np.empty_like(w), np.empty_like(w), np.zeros(3, dtype=np.float64)
    lib().orc_rootq_wt(_p(w), _p(dyv), _p(y), _p(dw), _p(gr), C.c_int64(w.size), C.c_float(U), C.c_float(L),
                       C.c_float(alpha), C.c_float(q))
    return y, dw, gr


def minmax(rows, n_bits, signed):
    rows = _f(rows)
    c, k = rows.shape
    s, o = np.empty(c, np.float32), np.empty(c, np.float32)
    lib().orc_minmax(_p(rows), C.c_int64(c), C.c_int64(k), C.c_int(n_bits), C.c_int(int(signed)), _p(s), _p(o))
    return s, o


def sweep_tensor(x, rows_for_mean, n_bits):
    x = _f(x).reshape(-1)
    s, o, losses = C.c_float(0), C.c_float(0), np.zeros(80, np.float64)
    pick = lib().orc_sweep_tensor(_p(x), C.c_int64(x.size), C.c_double(rows_for_mean), C.c_int(n_bits), C.byref(s),
                                  C.byref(o), _p(losses))
    return s.value, o.value, pick, losses


def sweep_channel(rows, n_bits, signed):
    rows = _f(rows)
    c, k = rows.shape
    s, o = np.empty(c, np.float32), np.empty(c, np.float32)
    lib().orc_sweep_channel(_p(rows), C.c_int64(c), C.c_int64(k), C.c_int(n_bits), C.c_int(int(signed)), _p(s), _p(o))
    return s, o


def l2norm_tensor(x, n_bits, signed, max_iters=1000):
    x = _f(x).reshape(-1)
    s, o = C.c_float(0), C.c_float(0)
    it = lib().orc_l2norm_tensor(_p(x), C.c_int64(x.size), C.c_int(n_bits), C.c_int(int(signed)), C.c_int(max_iters),
                                 C.byref(s), C.byref(o))
    return s.value, o.value, it


def merge_bn(w, bias, gamma, beta, mean, var):
    w = _f(w)
    c = w.shape[0]
    inner = w.size // c
    wo, bo = np.empty_like(w), np.empty(c, np.float32)
    b = None if bias is None else _f(bias)
    lib().orc_merge_bn(_p(w), _p(b), _p(_f(gamma)), _p(_f(beta)), _p(_f(mean)), _p(_f(var)), C.c_int64(c),
                       C.c_int64(inner), _p(wo), _p(bo))
    return wo, bo


def repvgg_fuse(k3, bn3, k1, bn1, bn_id, eps):
    """bn* = (gamma, beta, mean, var); bn_id None when the block has no identity branch."""
    k3, k1 = _f(k3), _f(k1)
    c, cin_g = k3.shape[0], k3.shape[1]
    wo, bo = np.empty_like(k3), np.empty(c, np.float32)
    a3 = [_f(t) for t in bn3]
    a1 = [_f(t) for t in bn1]
    ai = [None] * 4 if bn_id is None else [_f(t) for t in bn_id]
    lib().orc_repvgg_fuse(_p(k3), *[_p(t) for t in a3], C.c_float(eps), _p(k1), *[_p(t) for t in a1], C.c_float(eps),
                          *[_p(t) for t in ai], C.c_float(eps), C.c_int64(c), C.c_int64(cin_g), _p(wo), _p(bo))
    return wo, bo


def kth_value(x, k, abs_input=False):
    x = _f(x).reshape(-1)
    fn = lib().orc_kth_value
    fn.restype = C.c_float
    return fn(_p(x), C.c_int64(x.size), C.c_int64(int(k)), C.c_int(int(abs_input)))


def code_gemm(a_codes, w_codes, m_a, o_a, z_a, m_w, bias=None, relu=False):
    a, w = _f(a_codes), _f(w_codes)
    m, k = a.shape
    n = w.shape[0]
    mw = _f(m_w).reshape(-1)
    b = None if bias is None else _f(bias).reshape(-1)
    out = np.empty((m, n), dtype=np.float32)
    lib().orc_code_gemm(_p(a), _p(w), C.c_int64(m), C.c_int64(n), C.c_int64(k), C.c_float(float(m_a)),
                        C.c_float(float(o_a)), C.c_float(float(z_a)), _p(mw), C.c_int64(mw.size), _p(b),
                        C.c_int(int(relu)), _p(out))
    return out
