"""dlmc/quantization/scalar/RootQ/base.py: RootQBase (learnable clipping bounds with EMA, root-function
gradient estimator).  State names / shapes as in the reference: parameters in_scale [], wt_upper [],
wt_lower [], wt_alpha [] (init 1/4); buffers in_offset, wt_offset (None), in_run_upper [], in_run_scale [],
in_init_state [], wt_run_upper [], wt_run_lower [], wt_init_state []; attribute momentum.

forward() = [lazy init from one statistics pass] -> one-thread prepare kernel (EMA, gradient mix,
in-place running-buffer update, base.py:92-106,131-145) -> fused activation / weight kernels."""
import math
from abc import ABCMeta, abstractmethod

import torch
from torch.nn import Module

from ... import dist as qdist
from ... import functional as F
from ..utils import get_qrange
from .function import *  # noqa: F401,F403
from .function import RootQActFunction, RootQWeightFunction


class RootQBase(Module):
    __metaclass__ = ABCMeta

    def __init__(self, qconfig: dict = None):
        super(RootQBase, self).__init__()
        self.initialize(qconfig)

    def initialize(self, qconfig):
        """base.py:37-65."""
        self.qconfig = qconfig
        self.wt_min_val, self.wt_max_val = get_qrange(qconfig['weight']['args']['signed'],
                                                      qconfig['weight']['args']['n_bits'])
        self.in_min_val, self.in_max_val = get_qrange(qconfig['input']['args']['signed'],
                                                      qconfig['input']['args']['n_bits'])
        dev = self.weight.device if isinstance(getattr(self, 'weight', None), torch.Tensor) else None

        def scalar(v):
            return torch.tensor(float(v), dtype=torch.float32, device=dev)

        self.register_parameter('in_scale', torch.nn.Parameter(scalar(1.)))
        self.register_buffer('in_offset', None)
        self.register_buffer('in_run_upper', scalar(0.))
        self.register_buffer('in_run_scale', scalar(0.))
        self.register_buffer('in_init_state', scalar(0.))
        self.register_parameter('wt_upper', torch.nn.Parameter(scalar(2 ** 2 - 1)))
        self.register_parameter('wt_lower', torch.nn.Parameter(scalar((-1) * (2 ** 2))))
        self.register_parameter('wt_alpha', torch.nn.Parameter(scalar(1. / 4)))
        self.register_buffer('wt_offset', None)
        self.register_buffer('wt_run_upper', scalar(0.))
        self.register_buffer('wt_run_lower', scalar(0.))
        self.register_buffer('wt_init_state', scalar(0.))
        self.momentum = qconfig['momentum']
        self._host_init = {'in': None, 'wt': None}

    def reset_qparams(self):
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

    @abstractmethod
    def _forward_func(self, input, weight):
        raise NotImplementedError

    def forward(self, input):
        q = self.qconfig
        weight_q = self.weight
        if q['input']['enable']:
            if not self._ready('in', self.in_init_state):
                # base.py:80: (max - min) / (qmax - qmin), from one statistics pass (all-reduced under DDP)
                stats = qdist.sync_stats(F.obs_stats(input.detach()))
                span = torch.full((1,), float(self.in_max_val - self.in_min_val), device=input.device)
                in_scale = ((stats[0, 1] - stats[0, 0]).reshape(1) / span).reshape(())
                self.in_scale.data.copy_(in_scale)
                self.in_run_scale.data.copy_(in_scale)
                self.in_init_state.fill_(1)
                self._host_init['in'] = True
            g_i = 1 / math.sqrt(input.numel() * self.in_max_val)                       # base.py:93
            state = F.rootq_act_prepare(self.in_scale, self.in_run_scale, self.momentum, g_i, self.in_min_val,
                                        self.in_max_val, self.training)                  # base.py:95-101,105-106
            input = RootQActFunction.apply(input.contiguous(), self.in_scale, state)     # base.py:108-111
        if q['weight']['enable']:
            if not self._ready('wt', self.wt_init_state):
                # base.py:115-116: +-2*mean|w|*sqrt(qmax)
                stats = F.obs_stats(self.weight.detach())
                up = F.absmean_from_stats(stats, self.weight.numel(), 2.0, math.sqrt(self.wt_max_val), 1).reshape(())
                self.wt_upper.data.copy_(up)
                self.wt_lower.data.copy_(-up)
                self.wt_run_upper.data.copy_(up)
                self.wt_run_lower.data.copy_(-up)
                self.wt_init_state.fill_(1)
                self._host_init['wt'] = True
            g_w = 1 / math.sqrt(self.weight.numel() * self.wt_max_val)                  # base.py:136
            state = F.rootq_wt_prepare(self.wt_upper, self.wt_lower, self.wt_alpha, self.wt_run_upper,
                                       self.wt_run_lower, self.momentum, g_w, self.wt_min_val, self.wt_max_val,
                                       self.training)                                    # base.py:137-147
            weight_q = RootQWeightFunction.apply(self.weight.contiguous(), self.wt_upper, self.wt_lower,
                                                 self.wt_alpha, state)                   # base.py:146-155
        return self._forward_func(input, weight_q)
