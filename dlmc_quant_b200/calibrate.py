"""Weight-side calibration of a whole model in a constant number of launches.

The reference initialises every layer's weight quantizer lazily inside that layer's first forward
(modules/base.py:106-129, FSPTQuant/base.py:111-131): for the per-channel MSE clip search (`l2loss_channel`,
ops.py:169-196) that is one 80-candidate sweep launch per layer, 50-150 us each however small the layer - a quarter of
the calibration time of a RepVGG / MobileOne PTQ run.  The weights do not depend on the calibration data, so
`init_weight_quantizers(model)` computes ALL layers' weight (scale, offset) up front with one grouped sweep
(`dlmcq_obs_sweep_channel_grouped`) and marks the layers initialised; the first forward then only has the activation
observers left.  Values are bit-identical to the lazy per-layer initialisation (same kernel, same per-tensor geometry).

Layers whose weight observer is not a per-channel sweep over dim 0 are left to their lazy path."""
import torch

from . import functional as F
from .scalar.FSPTQuant.base import FSPTQBase
from .scalar.modules.base import QBase

__all__ = ["init_weight_quantizers"]


def _eligible(m):
    if not isinstance(m, (QBase, FSPTQBase)):
        return False
    q = m.qconfig["weight"]
    if not q["enable"] or q["type"] != "l2loss_channel" or q["args"].get("ch_axis", 0) != 0:
        return False
    if (getattr(m, "_host_init", None) or {}).get("wt"):
        return False
    return m.weight.is_cuda and m.weight.dtype in (torch.float32, torch.bfloat16)


def init_weight_quantizers(model):
    """Returns the number of layers initialised."""
    mods = [m for m in model.modules() if _eligible(m)]
    done = 0
    groups = {}
    for m in mods:
        a = m.qconfig["weight"]["args"]
        groups.setdefault((m.weight.device, m.weight.dtype, a["n_bits"], bool(a["signed"])), []).append(m)
    for (_, _, n_bits, signed), ms in groups.items():
        rows = [m.weight.detach().reshape(m.weight.shape[0], -1) for m in ms]
        for m, (scale, offset) in zip(ms, F.sweep_channel_grouped(rows, n_bits, signed)):
            shape = [m.weight.shape[0]] + [1] * (m.weight.dim() - 1)
            scale, offset = scale.reshape(shape), offset.reshape(shape)
            if isinstance(m, FSPTQBase):
                m.wt_offset = offset.detach().float().reshape(m.wt_scale.shape)
                m.wt_scale.data.copy_(scale.reshape(m.wt_scale.shape) + 1e-6)             # FSPTQuant/base.py:129
                if m.qconfig["weight"].get("recon_type") in ("adaround", "dist_recon"):
                    m.init_alpha()
            else:
                m.wt_offset = offset.detach().float()
                m._set_scale("wt_scale", scale)
            m.wt_init_state.fill_(1)
            m._host_init["wt"] = True
            done += 1
    return done
