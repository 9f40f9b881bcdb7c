"""Step-level grouping of the WEIGHT quantizers of a quantised model.

`QBase.forward` fake-quantises its weight per layer (modules/base.py:131-133): for ResNet-50 that is 54 small
forward launches and 54 x 2 backward launches per training step, all launch-latency-bound, plus ~45 us of host
work per call.  `group_weight_quantizers(model)` installs a forward pre-hook that quantises ALL (initialised)
weight tensors with ONE `dlmcq_fq_forward_grouped` launch before the model's forward, through ONE autograd node
whose backward is ONE `dlmcq_fq_backward_grouped` launch; each module then finds its quantised weight ready.
Same arithmetic as the per-layer path (same row kernels): w_q and dw bit-identical, scale gradients equal up to
summation order.  Modules whose weight quantizer is disabled, not yet initialised (first forward: the lazy
observer init runs per layer as usual) or not per-tensor / per-output-channel stay on the per-layer path."""
import ctypes as C
import math

import torch

from ... import _lib
from ... import functional as F
from ..._lib import FORM_AFFINE
from .base import QPARAMS_EPOCH, QBase

__all__ = ["WeightQuantGroup", "GroupHandle", "group_weight_quantizers"]


class _GroupedWeightFakeQuant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, n, *tensors):
        ctx.plan, ctx.n = plan, n             # the plan (membership + tables) this forward ran with
        ctx.save_for_backward(*tensors)
        return tuple(plan.forward(tensors[:n]))

    @staticmethod
    def backward(ctx, *grads):
        t = ctx.saved_tensors
        dws, dss = ctx.plan.backward(t[:ctx.n], t[ctx.n:], grads)
        return (None, None, *dws, *dss)


class _Plan:
    """Descriptor tables for one fixed membership (same device, same dtype).  Forward stores the plan in its
    autograd context, so backward uses the tables of the membership it was computed with even if the group
    re-plans in between (reset_qparams, a new layer becoming ready)."""

    def __init__(self, mods):
        self.mods = mods
        dev = mods[0].weight.device
        self.device, self.dtype = dev, mods[0].weight.dtype
        n = len(mods)
        self.arr = (_lib.GroupItem * n)()
        self.offsets = []                                       # float32 [channels] copies of the offset buffers
        units, chans, elems = [0], [0], [0]
        for i, m in enumerate(mods):
            w, s = m.weight, m.wt_scale
            c = s.numel()
            off = m.wt_offset.detach().to(device=dev, dtype=torch.float32).reshape(-1)   # a view when already fp32
            off = (off.expand(c) if off.numel() == 1 and c > 1 else off).contiguous()
            self.offsets.append(off)
            it = self.arr[i]
            it.x, it.scale, it.offset = w.data_ptr(), s.data_ptr(), off.data_ptr()
            it.channels, it.inner = c, w.numel() // c
            it.form, it.lo, it.hi = FORM_AFFINE, int(m.wt_min_val), int(m.wt_max_val)
            it.g = 1 / math.sqrt(w.numel() * m.wt_max_val)                      # base.py:131
            units.append(units[-1] + c * ((it.inner + F.GROUP_SEG - 1) // F.GROUP_SEG))
            chans.append(chans[-1] + c)
            elems.append(elems[-1] + w.numel())
        self.units, self.chans, self.elems = units, chans, elems
        self.unit_prefix = torch.tensor(units, dtype=torch.int64).to(dev)
        self.chan_prefix = torch.tensor(chans, dtype=torch.int64).to(dev)
        self.partials = torch.empty(max(units[-1], 1), dtype=torch.float32, device=dev)
        self.es = mods[0].weight.element_size()
        # descriptor upload without a stream synchronisation: a ring of pinned staging buffers (a blocking
        # .to(device) from pageable memory would drain the GPU twice per step)
        nbytes = C.sizeof(self.arr)
        self._ring = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
        self._ring_ev = [None] * 4
        self._ring_i = 0
        self._table = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def _upload(self):
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture: the copy node re-reads its pinned source on every replay, so each captured upload
            # gets a staging buffer of its own that is never rewritten (the pointers in it - graph-pool allocations
            # and parameters - are the same on every replay), and no event may be waited on while capturing
            pin = torch.empty(C.sizeof(self.arr), dtype=torch.uint8).pin_memory()
            self._captured = getattr(self, "_captured", [])
            self._captured.append(pin)
            C.memmove(pin.data_ptr(), C.addressof(self.arr), pin.numel())
            self._table.copy_(pin, non_blocking=True)
            return self._table
        k = self._ring_i
        self._ring_i = (k + 1) % len(self._ring)
        if self._ring_ev[k] is None:
            self._ring_ev[k] = torch.cuda.Event()
        else:
            self._ring_ev[k].synchronize()                      # four uploads ago: long complete
        pin = self._ring[k]
        C.memmove(pin.data_ptr(), C.addressof(self.arr), pin.numel())
        self._table.copy_(pin, non_blocking=True)               # stream-ordered before the launch that reads it
        self._ring_ev[k].record()
        return self._table

    def _like(self, flat, i, w):
        return torch.as_strided(flat, w.shape, w.stride(), storage_offset=self.elems[i])

    # -- the two launches -----------------------------------------------------------------------------------
    def forward(self, weights):
        flat = torch.empty(self.elems[-1], dtype=self.dtype, device=self.device)
        base = flat.data_ptr()
        for i in range(len(weights)):
            self.arr[i].y = base + self.elems[i] * self.es
            self.arr[i].dy = None
            self.arr[i].dscale = None
        items = self._upload()
        with F._on(self.device):
            _lib.check(_lib.lib().dlmcq_fq_forward_grouped(items.data_ptr(), self.unit_prefix.data_ptr(), len(weights),
                                                           self.units[-1], F._dtype_code(flat), F._stream_ptr()))
        return [self._like(flat, i, w) for i, w in enumerate(weights)]

    def backward(self, weights, scales, grads):
        flat = torch.empty(self.elems[-1], dtype=self.dtype, device=self.device)
        ds = torch.empty(self.chans[-1], dtype=torch.float32, device=self.device)
        base, dsb = flat.data_ptr(), ds.data_ptr()
        keep = []
        for i, w in enumerate(weights):
            g = grads[i]
            if g is None:
                g = torch.zeros_like(w)
            elif g.stride() != w.stride() or g.dtype != w.dtype:
                g = torch.empty_like(w).copy_(g)                # same element order as the weight
            keep.append(g)
            it = self.arr[i]
            it.dy, it.y, it.dscale = g.data_ptr(), base + self.elems[i] * self.es, dsb + 4 * self.chans[i]
        items = self._upload()
        with F._on(self.device):
            _lib.check(_lib.lib().dlmcq_fq_backward_grouped(items.data_ptr(), self.unit_prefix.data_ptr(),
                                                            self.chan_prefix.data_ptr(), len(weights), self.units[-1],
                                                            self.chans[-1], F._dtype_code(flat),
                                                            self.partials.data_ptr(), F._stream_ptr()))
        dws = [self._like(flat, i, w) for i, w in enumerate(weights)]
        dss = [ds[self.chans[i]:self.chans[i + 1]].reshape(s.shape).to(s.dtype) for i, s in enumerate(scales)]
        return dws, dss


class WeightQuantGroup:
    def __init__(self, model, modules=None):
        self.model = model
        self._key = None
        self._plan = None
        self._epoch = -1
        self._mods = []
        self._candidates = list(modules) if modules is not None else None

    # -- membership ---------------------------------------------------------------------------------------
    @staticmethod
    def _eligible(m):
        if not isinstance(m, QBase) or not m.qconfig['weight']['enable']:
            return False
        if not (getattr(m, '_host_init', None) or {}).get('wt'):
            return False                                       # observer init pending (first forward / reset_qparams)
        w, s = m.weight, m.wt_scale
        if not w.is_cuda or w.dtype not in (torch.float32, torch.bfloat16) or m.wt_offset is None:
            return False
        per_channel = s.numel() == w.shape[0] and s.numel() > 1 and tuple(s.shape[1:]) == (1,) * (w.dim() - 1)
        return (s.numel() == 1 or per_channel) and F.dense_as_is(w, 0)

    @staticmethod
    def _offset_in_place(m):
        o = m.wt_offset
        return o.dtype is torch.float32 and o.is_contiguous() and o.numel() == m.wt_scale.numel() and \
            o.device == m.weight.device

    def _refresh(self):
        if self._candidates is None:
            self._candidates = [m for m in self.model.modules() if isinstance(m, QBase)]
        # membership is re-evaluated while layers are still missing and whenever ANY layer's quantizer state was
        # invalidated since the last look (reset_qparams - the reference's QATTrainer calls it every
        # update_qparams_period steps, qat_trainer.py:44-48 - or a checkpoint load): such a layer leaves the
        # group, re-observes its weight on the per-layer path in this forward, and rejoins on the next one
        if len(self._mods) < len(self._candidates) or self._epoch != QPARAMS_EPOCH[0]:
            self._epoch = QPARAMS_EPOCH[0]
            mods = [m for m in self._candidates if self._eligible(m)]
            if mods:        # one launch = one device and one element type; the rest stays on the per-layer path
                dev, dt = mods[0].weight.device, mods[0].weight.dtype
                mods = [m for m in mods if m.weight.device == dev and m.weight.dtype == dt]
        else:
            mods = self._mods
        # An in-place update of an offset buffer (load_state_dict; DistributedDataParallel re-broadcasts every buffer on
        # every forward) must refresh the plan only when the plan holds a converted COPY of it; a float32 buffer with
        # one entry per channel is used in place (the plan's `off` is a view of it), so its _version is irrelevant -
        # keying on it rebuilt the whole plan on every DDP step
        key = tuple((id(m), m.weight.data_ptr(), m.wt_scale.data_ptr(), m.wt_offset.data_ptr(),
                     0 if self._offset_in_place(m) else m.wt_offset._version, m.weight.stride(), m.weight.dtype)
                    for m in mods)
        if key == self._key:
            return
        self._key, self._mods = key, mods
        self._plan = _Plan(mods) if mods else None

    # -- hook --------------------------------------------------------------------------------------------------
    def __call__(self, module=None, args=None):
        self._refresh()
        if not self._mods:
            return None
        n = len(self._mods)
        outs = _GroupedWeightFakeQuant.apply(self._plan, n, *[m.weight for m in self._mods],
                                             *[m.wt_scale for m in self._mods])
        for m, o in zip(self._mods, outs):
            m.__dict__['_wq'] = o                                # consumed (popped) by QBase.forward
        return None


class GroupHandle:
    """Returned by group_weight_quantizers: remove() restores the per-layer behaviour."""

    def __init__(self, model, groups, hook):
        self.model, self.groups, self._hook = model, groups, hook
        self.group = groups[0]

    def remove(self):
        self._hook.remove()
        for m in self.model.modules():
            m.__dict__.pop('_wq', None)


def group_weight_quantizers(model, n_groups=None):
    """Quantise all weight tensors of `model`'s QBase layers in one launch per direction (see module docstring).

    n_groups > 1 splits the layers, in forward order, into that many groups of about equal parameter bytes, each with
    its own autograd node: a group's backward runs as soon as ITS layers' weight gradients exist, so a data-parallel
    wrapper can overlap the all-reduce of the later layers' gradients with the rest of the backward pass (one node
    for the whole model delivers every weight gradient at the very end).  Default: 1, or 4 when torch.distributed
    is initialised with more than one rank."""
    if n_groups is None:
        import torch.distributed as dist
        n_groups = 4 if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else 1
    mods = [m for m in model.modules() if isinstance(m, QBase)]
    n_groups = max(1, min(int(n_groups), len(mods) or 1))
    total = sum(m.weight.numel() for m in mods) or 1
    chunks, cur, acc = [], [], 0
    for m in mods:
        cur.append(m)
        acc += m.weight.numel()
        if acc >= total * (len(chunks) + 1) / n_groups and len(chunks) < n_groups - 1:
            chunks.append(cur)
            cur = []
    if cur:
        chunks.append(cur)
    groups = [WeightQuantGroup(model, c) for c in chunks]

    def hook(module, args):
        for g in groups:
            g(module, args)
        return None
    return GroupHandle(model, groups, model.register_forward_pre_hook(hook))
