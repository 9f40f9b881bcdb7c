"""dlmc/quantization/scalar/FSPTQuant/base.py: FSPTQBase (RepAPQ / FSPTQ post-training quantisation with
per-channel weight scales and optional AdaRound soft rounding).

State as in the reference: parameters in_scale [1], wt_scale [C,1,1,1] or [C,1], optional alpha (like
weight); buffers in_offset [1], wt_offset (like wt_scale), in_init_state [1], wt_init_state [1],
org_weight; attributes act_quant, wt_quant, soft_target, gamma, zeta, beta, train_module.

Deliberate differences: buffers are created on the weight's device instead of a hard-coded
torch.device('cuda') (base.py:47); with act_quant off the raw input is used (the reference leaves
q_input unbound, base.py:97,154-156); the dead `dist_recon` branch (base.py:133,143 call an undefined
method and torch.floor() without arguments) is treated like 'adaround'."""
from abc import ABCMeta, abstractmethod

import torch
from torch.nn import Module

from ... import functional as F
from ..._lib import FORM_SYM, FORM_ZP
from ..modules.function import *  # noqa: F401,F403
from ..modules.function import FakeQuantFunction, fake_quantize
from ..ops import get_qparams_tensor
from ..utils import get_qrange


class AdaRoundFunction(torch.autograd.Function):
    """base.py:136-141,151-152 in training mode: floor(w/s) + h(alpha), clamp, * s.  torch.floor has no
    straight-through in the reference, so d/dw is zero; the gradients are d(scale) and d(alpha)."""

    @staticmethod
    def forward(ctx, w, scale, alpha, lo, hi):
        ctx.save_for_backward(w, scale, alpha)
        ctx.cfg = (lo, hi)
        return F.adaround_forward(w, alpha, scale, lo, hi, soft=True)

    @staticmethod
    def backward(ctx, dy):
        w, scale, alpha = ctx.saved_tensors
        lo, hi = ctx.cfg
        dalpha, ds = F.adaround_backward(w, alpha, dy, scale, lo, hi)
        return torch.zeros_like(w), ds.reshape(scale.shape), dalpha, None, None


class FSPTQBase(Module):
    __metaclass__ = ABCMeta

    def __init__(self, qconfig: dict = None):
        super(FSPTQBase, self).__init__()
        self.initialize(qconfig)

    def initialize(self, qconfig):
        """base.py:33-63."""
        self.qconfig = qconfig
        self.train_module = 0
        self.wt_min_val, self.wt_max_val = get_qrange(qconfig['weight']['args']['signed'],
                                                      qconfig['weight']['args']['n_bits'])
        self.in_min_val, self.in_max_val = get_qrange(qconfig['input']['args']['signed'],
                                                      qconfig['input']['args']['n_bits'])
        dev = self.weight.device
        channel = self.weight.shape[0]
        shape = (channel, 1, 1, 1) if self.weight.dim() == 4 else (channel, 1)
        self.register_parameter('in_scale', torch.nn.Parameter(torch.ones(1, device=dev)))
        self.register_buffer('in_offset', torch.zeros(1, device=dev))
        self.register_buffer('in_init_state', torch.zeros(1, device=dev))
        self.register_parameter('wt_scale', torch.nn.Parameter(torch.ones(shape, device=dev)))
        self.register_buffer('wt_offset', torch.ones(shape, device=dev))
        self.register_buffer('wt_init_state', torch.zeros(1, device=dev))
        self.register_buffer('org_weight', self.weight.clone().detach())
        self.act_quant = self.qconfig['input']['enable']
        self.wt_quant = self.qconfig['weight']['enable']
        self.soft_target = True
        if self.qconfig['weight'].get('recon_type') in ["adaround", "dist_recon"]:
            self.register_parameter('alpha', torch.nn.Parameter(torch.ones_like(self.weight)))
            self.gamma, self.zeta = -0.1, 1.1
            self.beta = 2 / 3
        self._host_init = {'in': None, 'wt': None}

    @abstractmethod
    def _forward_func(self, input, weight):
        raise NotImplementedError

    def init_alpha(self):
        """base.py:69-76."""
        self.alpha.data.copy_(F.adaround_init_alpha(self.weight.detach(), self.wt_scale.detach()))

    def get_soft_targets(self):
        """base.py:78-79 (tiny composite, used by the trainer's regulariser, not by forward)."""
        return torch.clamp(torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def change_quant_state(self, wt_state, act_state):
        self.wt_quant = wt_state
        self.act_quant = act_state

    def reinit_parameters(self):
        self.in_init_state.fill_(0)
        self.wt_init_state.fill_(0)
        self._host_init = {'in': False, 'wt': False}

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._host_init = {'in': None, 'wt': None}

    def _ready(self, which, flag):
        h = getattr(self, '_host_init', None)
        if h is None:
            h = self._host_init = {'in': None, 'wt': None}
        if h[which] is None:
            h[which] = bool(flag.item() != 0)
        return h[which]

    def forward(self, input):
        q = self.qconfig
        q_input = input
        if self.act_quant:
            if not self._ready('in', self.in_init_state):
                scale, offset = get_qparams_tensor(input.detach(), qtype=q['input']['type'], **q['input']['args'])
                self.in_offset = offset.detach().float().reshape(-1)[:1].clone() if offset.numel() == 1 else offset.detach().float()
                self.in_scale.data.copy_(scale.reshape(self.in_scale.shape))
                self.in_init_state.fill_(1)
                self._host_init['in'] = True
            q_input = fake_quantize(input, self.in_scale, self.in_offset, self.in_min_val, self.in_max_val,
                                    FORM_ZP)                                              # base.py:108-109
        if not self.wt_quant:
            return self._forward_func(q_input, self.weight)
        recon = q['weight'].get('recon_type')
        if not self._ready('wt', self.wt_init_state):
            scale, offset = get_qparams_tensor(self.weight.detach(), qtype=q['weight']['type'], **q['weight']['args'])
            self.wt_offset = offset.detach().float().reshape(self.wt_scale.shape)
            self.wt_scale.data.copy_(scale.reshape(self.wt_scale.shape) + 1e-6)            # base.py:129
            if recon in ('adaround', 'dist_recon'):
                self.init_alpha()
            self.wt_init_state.fill_(1)
            self._host_init['wt'] = True
        if recon in ('adaround', 'dist_recon'):
            if self.training:
                weight = AdaRoundFunction.apply(self.weight.contiguous(), self.wt_scale, self.alpha, self.wt_min_val,
                                                self.wt_max_val)                           # base.py:137-139
            else:
                weight = F.adaround_forward(self.weight, self.alpha, self.wt_scale, self.wt_min_val,
                                            self.wt_max_val, soft=False)                   # base.py:141
        else:
            weight = FakeQuantFunction.apply(self.weight.contiguous(), self.wt_scale, None, self.wt_min_val,
                                             self.wt_max_val, FORM_SYM, 0.0, 0)            # base.py:149-152
        return self._forward_func(q_input, weight)
