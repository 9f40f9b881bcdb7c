"""Shared by the CPU and GPU tests of the integer-code layer product: turns a fixture of tests/golden/qgemm.npz (minted
by oracle/make_golden.py::golden_qgemm from the UNMODIFIED reference: the eval-mode output `y` of its QLinear / 1x1 QConv2d
/ FSPTQ modules, the fake-quantised tensors it fed to `_forward_func`, and the qparams its observers chose) into the
matrices and quantizer descriptions the factored product needs."""
import torch

from oracle import restate as R

A1, AFFINE, ZP, SYM = 0, 1, 2, 3


def _rows(t):
    """[B, C, H, W] -> [B*H*W, C] (the channels-last activation matrix); 2-D tensors pass through."""
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous() if t.dim() == 4 else t.contiguous()


class Problem:
    def __init__(self, case):
        q = case.meta["qconfig"]
        self.name, self.family = case.name, case.meta["family"]
        self.a_lo, self.a_hi = R.qrange(q["input"]["args"]["signed"], q["input"]["args"]["n_bits"])
        self.w_lo, self.w_hi = R.qrange(q["weight"]["args"]["signed"], q["weight"]["args"]["n_bits"])
        x, w = case.inp["x"], case.inp["weight"]
        self.x, self.w = _rows(x), w.reshape(w.shape[0], -1).contiguous()
        self.bias = case.inp["bias"] if case.inp["bias"].numel() else None
        self.y, self.qx, self.qw = _rows(case.out["y"]), _rows(case.out["qx"]), case.out["qw"].reshape(w.shape[0], -1)
        self.s_a = case.out["param_in_scale"].reshape(1)
        self.off_a = case.out["buf_in_offset"].reshape(1)
        self.s_w = case.out["param_wt_scale"].reshape(-1)
        assert not bool(case.out["buf_wt_offset"].any()), "fixture premise: symmetric weights"
        if self.family == "qbase":          # modules/base.py:96-102,131-133
            self.a_form = self.w_form = AFFINE
            self.g_a, self.g_w = R.lsq_g(x.numel(), self.a_hi), R.lsq_g(w.numel(), self.w_hi)
        else:                               # FSPTQuant/base.py:108-109,149-152
            self.a_form, self.w_form, self.g_a, self.g_w = ZP, SYM, 0.0, 0.0

    def oracle(self):
        """(activation codes, weight codes, m_a, o_a, z_a, m_w) through the oracle's restatement of the two quantizers."""
        sw = self.s_w.reshape(-1, 1) if self.s_w.numel() > 1 else self.s_w
        if self.a_form == AFFINE:
            ca = R.fq_affine_codes(self.x, self.s_a, self.off_a, self.a_lo, self.a_hi, self.g_a)
            cw = R.fq_affine_codes(self.w, sw, torch.zeros(1), self.w_lo, self.w_hi, self.g_w)
            return (ca.detach(), cw.detach(), R.grad_scale(self.s_a, self.g_a).detach(), self.off_a, torch.zeros(1),
                    R.grad_scale(self.s_w, self.g_w).detach())
        ca = R.fq_zp_codes(self.x, self.s_a, self.off_a, self.a_lo, self.a_hi)
        cw = R.fq_sym_codes(self.w, sw, self.w_lo, self.w_hi)
        return ca.detach(), cw.detach(), self.s_a, torch.zeros(1), self.off_a, self.s_w

    def bound(self):
        """sum_k |y_a * y_w| (+ |bias|): the scale of the 1e-5 tolerance."""
        mag = self.qx.double().abs() @ self.qw.double().abs().t()
        return mag if self.bias is None else mag + self.bias.double().abs()
