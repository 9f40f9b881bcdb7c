"""dlmc/quantization/scalar/modules/base.py: QBase, the generic QAT / PTQ / LSQ quantised layer.

Same state names, shapes and lazy-init behaviour as the reference (checkpoints and the trainers'
fnmatch filters depend on them): parameters in_scale [1], wt_scale [1]; buffers in_offset, wt_offset
(None until first forward), in_init_state [1], wt_init_state [1]; ints in/wt_min_val, in/wt_max_val.
What changes is the arithmetic inside forward(): the 8-op eager chain per tensor (base.py:96-102,
131-133) and its autograd become one fused forward and one fused backward kernel each.

Deliberate differences: no device-to-host sync per forward (the reference tests a device
buffer `in_init_state == 0` every call, base.py:82,107 - here a host flag mirrors it), no hard-coded
torch.device('cuda') (base.py:85,120), per-channel scales are allocated to fit instead of crashing in
copy_ (base.py:116,128), observer statistics are all-reduced across DDP ranks."""
import math
from abc import ABCMeta, abstractmethod
from fnmatch import fnmatch

import torch
from torch.nn import Module

from ... import functional as F
from ..._lib import FORM_AFFINE
from ..ops import get_qparams_output, get_qparams_tensor
from ..utils import get_qrange
from .function import *  # noqa: F401,F403  (the reference re-exports the Functions from base)
from .function import fake_quantize
from ... import dist as qdist


class QBase(Module):
    __metaclass__ = ABCMeta

    qconfig: dict
    wt_min_val: int
    wt_max_val: int
    wt_scale: torch.nn.Parameter
    in_min_val: int
    in_max_val: int
    in_scale: torch.nn.Parameter

    def __init__(self, qconfig: dict = None):
        super(QBase, self).__init__()
        self.initialize(qconfig)

    def initialize(self, qconfig):
        """base.py:28-55."""
        if 'channel' in str(qconfig['input']['type']):
            qconfig['input']['args']['ch_axis'] = 1
        self.qconfig = qconfig
        self.wt_min_val, self.wt_max_val = get_qrange(qconfig['weight']['args']['signed'],
                                                      qconfig['weight']['args']['n_bits'])
        self.in_min_val, self.in_max_val = get_qrange(qconfig['input']['args']['signed'],
                                                      qconfig['input']['args']['n_bits'])
        dev = self.weight.device if isinstance(getattr(self, 'weight', None), torch.Tensor) else None
        self.register_parameter('in_scale', torch.nn.Parameter(torch.ones(1, device=dev)))
        self.register_buffer('in_offset', None)
        self.register_buffer('in_init_state', torch.zeros(1, device=dev))
        self.register_parameter('wt_scale', torch.nn.Parameter(torch.ones(1, device=dev)))
        self.register_buffer('wt_offset', None)
        self.register_buffer('wt_init_state', torch.zeros(1, device=dev))
        self._host_init = {'in': None, 'wt': None}      # host mirror of *_init_state (None = unknown)

    def reset_qparams(self):
        """base.py:57-61 sets the scales to None (and then crashes in grad_scale); here the intent -
        re-run the observers on the next forward - is implemented instead."""
        self.in_init_state.fill_(0)
        self.wt_init_state.fill_(0)
        self._host_init = {'in': False, 'wt': False}

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._host_init = {'in': None, 'wt': None}       # re-read the flags after a checkpoint load

    def _ready(self, which, flag):
        h = getattr(self, '_host_init', None)
        if h is None:
            h = self._host_init = {'in': None, 'wt': None}
        if h[which] is None:
            h[which] = bool(flag.item() != 0)            # one sync, first call only
        return h[which]

    def _set_scale(self, name, value):
        p = getattr(self, name)
        value = value.detach().to(p.device, torch.float32)
        if value.numel() == p.numel():
            p.data.copy_(value.reshape(p.shape))
        else:                                            # per-channel observer: allocate to fit
            setattr(self, name, torch.nn.Parameter(value.clone()))

    @abstractmethod
    def _forward_func(self, input, weight):
        raise NotImplementedError

    def _lsq_init(self, t, qmax):
        """base.py:84,119: 2*mean|x|/sqrt(qmax), offset zeros(1)."""
        stats = qdist.sync_stats(F.obs_stats(t.detach()))
        count = t.numel() * qdist.world_size()
        return F.absmean_from_stats(stats, count, 2.0, math.sqrt(qmax), 0), torch.zeros(1, device=t.device)

    def forward(self, input):
        q = self.qconfig
        if q['input']['enable']:
            if not self._ready('in', self.in_init_state):
                if fnmatch(q['input']['type'], 'LSQ'):
                    scale, self.in_offset = self._lsq_init(input, self.in_max_val)
                else:
                    scale, offset = get_qparams_tensor(input.detach(), qtype=q['input']['type'], **q['input']['args'])
                    self.in_offset = offset.detach().float()
                self._set_scale('in_scale', scale)
                self.in_init_state.fill_(1)
                self._host_init['in'] = True
            g_i = 1 / math.sqrt(input.numel() * self.in_max_val)                       # base.py:96
            input = fake_quantize(input, self.in_scale, self.in_offset, self.in_min_val, self.in_max_val,
                                  FORM_AFFINE, g_i)                                      # base.py:97,102
        weight = self.weight
        grouped = self.__dict__.pop('_wq', None)         # set by group_weight_quantizers' pre-hook for this step
        if grouped is not None and q['weight']['enable']:
            weight = grouped
        elif q['weight']['enable']:
            if not self._ready('wt', self.wt_init_state):
                if fnmatch(q['weight']['type'], '*output*'):
                    scale, offset = get_qparams_output(input.detach(), self.weight.detach(), self,
                                                       qtype=q['weight']['type'], **q['weight']['args'])
                    self.wt_offset = offset.detach().float()
                elif fnmatch(q['weight']['type'], 'LSQ'):
                    scale, self.wt_offset = self._lsq_init(self.weight, self.wt_max_val)
                else:
                    scale, offset = get_qparams_tensor(self.weight.detach(), qtype=q['weight']['type'],
                                                       **q['weight']['args'])
                    self.wt_offset = offset.detach().float()
                self._set_scale('wt_scale', scale)
                self.wt_init_state.fill_(1)
                self._host_init['wt'] = True
            g_w = 1 / math.sqrt(self.weight.numel() * self.wt_max_val)                  # base.py:131
            weight = fake_quantize(self.weight, self.wt_scale, self.wt_offset, self.wt_min_val, self.wt_max_val,
                                   FORM_AFFINE, g_w)                                     # base.py:132-133
        return self._forward_func(input, weight)
