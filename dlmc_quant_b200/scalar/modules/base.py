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

# bumped whenever any layer's quantizer state is invalidated (reset_qparams, load_state_dict): step-level caches
# (modules/group.py) re-evaluate their membership when it changes
QPARAMS_EPOCH = [0]


def _channel_scale_shape(weight, qtype, axis_of_weight, lead=0):
    """Shape of a per-channel scale known at initialize() time (so that the Parameter object an optimizer / DDP
    reducer captured before the first forward is the one the observer fills), or None for per-tensor types."""
    if 'channel' not in str(qtype) or not isinstance(weight, torch.Tensor):
        return None
    return [1] * lead + [weight.shape[axis_of_weight]] + [1] * (weight.dim() - 1 - lead)


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
        w = getattr(self, 'weight', None)
        dev = w.device if isinstance(w, torch.Tensor) else None
        # per-channel types: the scale Parameters are sized here ([1,C,1,1] inputs of an ungrouped layer / [C,1,1,1]
        # weights with ch_axis 0), so they are never replaced after an optimizer or DDP captured them; the reference
        # registers [1] and then crashes in copy_ (base.py:116,128)
        in_shape = wt_shape = None
        if getattr(self, 'groups', 1) == 1:
            in_shape = _channel_scale_shape(w, qconfig['input']['type'], 1, lead=1)
        if qconfig['weight']['args'].get('ch_axis', 0) == 0 and 'output' not in str(qconfig['weight']['type']):
            wt_shape = _channel_scale_shape(w, qconfig['weight']['type'], 0)
        self.register_parameter('in_scale', torch.nn.Parameter(torch.ones(in_shape or 1, device=dev)))
        self.register_buffer('in_offset', None)
        self.register_buffer('in_init_state', torch.zeros(1, device=dev))
        self.register_parameter('wt_scale', torch.nn.Parameter(torch.ones(wt_shape or 1, device=dev)))
        self.register_buffer('wt_offset', None)
        self.register_buffer('wt_init_state', torch.zeros(1, device=dev))
        self._host_init = {'in': None, 'wt': None}      # host mirror of *_init_state (None = unknown)

    def reset_qparams(self):
        """base.py:57-61 sets the scales to None (and then crashes in grad_scale); here the intent -
        re-run the observers on the next forward - is implemented instead."""
        self.in_init_state.fill_(0)
        self.wt_init_state.fill_(0)
        self._host_init = {'in': False, 'wt': False}
        QPARAMS_EPOCH[0] += 1

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # a checkpoint written after calibration may hold per-channel scales / offsets where this (fresh) module
        # still has the [1] placeholders: take the checkpoint's shapes instead of failing the strict size check
        for name in ('in_scale', 'wt_scale'):
            v = state_dict.get(prefix + name)
            p = getattr(self, name, None)
            if isinstance(v, torch.Tensor) and p is not None and v.shape != p.shape:
                setattr(self, name, torch.nn.Parameter(torch.empty(v.shape, dtype=p.dtype, device=p.device)))
        for name in ('in_offset', 'wt_offset'):
            v = state_dict.get(prefix + name)
            b = getattr(self, name, None)
            if isinstance(v, torch.Tensor) and (b is None or b.shape != v.shape):
                setattr(self, name, torch.empty(v.shape, dtype=v.dtype, device=self.in_init_state.device))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._host_init = {'in': None, 'wt': None}       # re-read the flags after a checkpoint load
        QPARAMS_EPOCH[0] += 1

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
        else:
            # observer result of a shape initialize() could not foresee (pixel / grouped-input types): allocate
            # to fit.  An optimizer / DDP wrapper built BEFORE this first forward still holds the old Parameter -
            # calibrate (one forward) before building them for such types.
            setattr(self, name, torch.nn.Parameter(value.clone()))
            QPARAMS_EPOCH[0] += 1

    @abstractmethod
    def _forward_func(self, input, weight):
        raise NotImplementedError

    def _lsq_init(self, t, qmax):
        """base.py:84,119: 2*mean|x|/sqrt(qmax), offset zeros(1)."""
        stats = qdist.sync_stats(F.obs_stats(t.detach()))
        count = t.numel() * qdist.world_size()
        return F.absmean_from_stats(stats, count, 2.0, math.sqrt(qmax), 0), torch.zeros(1, device=t.device)

    def _quant_input(self, input):
        """base.py:82-102: lazy observer init, then the fused input fake-quant."""
        q = self.qconfig
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
        return fake_quantize(input, self.in_scale, self.in_offset, self.in_min_val, self.in_max_val,
                             FORM_AFFINE, g_i)                                      # base.py:97,102

    def _quant_weight(self, input):
        """base.py:106-133 (`input` is only used by the output-aware observers)."""
        q = self.qconfig
        # set by group_weight_quantizers' pre-hook for this step (membership test first: no dict mutation on the
        # common path, which also keeps the method traceable by torch.compile)
        grouped = self.__dict__.pop('_wq') if '_wq' in self.__dict__ else None
        if not q['weight']['enable']:
            return self.weight
        if grouped is not None:
            return grouped
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
        return fake_quantize(self.weight, self.wt_scale, self.wt_offset, self.wt_min_val, self.wt_max_val,
                             FORM_AFFINE, g_w)                                      # base.py:132-133

    def forward_prequantized(self, input_q):
        """The layer on an input that its producer already fake-quantised with THIS layer's in_scale / in_offset
        (dlmc_quant_b200.fuse: BatchNorm + ReLU + this quantizer run as one kernel)."""
        return self._forward_func(input_q, self._quant_weight(input_q))

    def forward(self, input):
        if self.qconfig['input']['enable']:
            pre = getattr(input, '_dlmcq_q', None)       # (consumer, a_q) attached by a fused producer (fuse.py)
            if pre is not None and pre[0] is self:
                input = pre[1]
            else:
                input = self._quant_input(input)
        return self._forward_func(input, self._quant_weight(input))
