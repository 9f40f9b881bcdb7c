"""dlmc/utils/quantize.py::quantize_model - the plug-in boundary: swaps nn.Conv2d / nn.Linear instances
for the quantised classes in place (class swap via __new__ + __dict__.update + initialize, exactly like
quantize.py:130-136 - __init__ is never called), with the same config schema (`weight` / `input` blocks,
`exclude_layers` regexes, `override_options`, `momentum`).  Host-side string plumbing only.

BitMixer / MetaQ mappings are not offered: those packages are absent from the reference tree as well
(quantize.py:10,12 import modules that do not exist)."""
import copy
import re
from operator import attrgetter, itemgetter
from typing import Dict, Iterable, List

from torch import nn

from .scalar import FSPTQuant as FSPQ
from .scalar import RootQ as RQ
from .scalar import modules as qnn
from .scalar.modules.group import group_weight_quantizers  # noqa: F401  (step-level weight grouping)

__all__ = ['quantize_model', 'get_layers', 'attrsetter', 'group_weight_quantizers']

MODULE_MAPPING = {nn.Conv2d: qnn.QConv2d, nn.Linear: qnn.QLinear}
ROOTQ_MAPPING = {nn.Conv2d: RQ.RootQConv2d, nn.Linear: RQ.RootQLinear}
FSPTQUANT_MAPPING = {nn.Conv2d: FSPQ.FSPTQConv2d, nn.Linear: FSPQ.FSPTQLinear}


def attrsetter(*items):
    """dlmc/utils/access.py:12-27."""
    def resolve_attr(obj, attr):
        attrs = attr.split(".")
        for name in attrs[:-1]:
            obj = getattr(obj, name)
        return obj, attrs[-1]

    def g(obj, val):
        for attr in items:
            resolved_obj, resolved_attr = resolve_attr(obj, attr)
            setattr(resolved_obj, resolved_attr, val)
    return g


def get_layers(model: nn.Module, filter_regexp: str = "(.*?)", filter_types: Iterable = None) -> List[str]:
    """dlmc/utils/access.py:30-61: layer names from parameter names, filtered by regex and type."""
    names = map(itemgetter(0), model.named_parameters())
    names = filter(lambda x: "bias" not in x, names)
    names = map(lambda x: x.replace(".weight_orig", ""), names)
    names = map(lambda x: x.replace(".weight", ""), names)
    r = re.compile("(module\\.)?" + "(" + filter_regexp + ")")
    names = list(filter(r.match, names))
    if filter_types is not None:
        names = [n for n in names if isinstance(attrgetter(n)(model), filter_types)]
    return names


def _override_options(dst_config: Dict, src_config: Dict = None) -> Dict:
    """quantize.py:44-58."""
    if src_config is None:
        return dst_config
    dst_config = copy.deepcopy(dst_config)
    if 'type' in src_config:
        dst_config['type'] = src_config['type']
    if 'enable' in src_config:
        dst_config['enable'] = src_config['enable']
    if 'args' in src_config:
        dst_config['args'].update(src_config['args'])
    return dst_config


def quantize_model(model: nn.Module, config: Dict, logger=None, quantization_type: str = None, **kwargs) -> None:
    """quantize.py:61-142."""
    default_weight_config = config['weight']
    default_input_config = config['input']
    default_momentum_config = 0.1
    if quantization_type == "RootQ":
        mapping = ROOTQ_MAPPING
        default_momentum_config = config['momentum']
    elif quantization_type == "FSPTQ":
        mapping = FSPTQUANT_MAPPING
    elif quantization_type in ("BitMixer", "MetaQ"):
        raise NotImplementedError(f"{quantization_type}: its package is not part of the reference tree either")
    else:
        mapping = MODULE_MAPPING

    all_layers = get_layers(model, filter_types=tuple(mapping.keys()))
    exclude_layers = []
    for regexp in config.get('exclude_layers', []):
        exclude_layers.extend(get_layers(model, filter_regexp=regexp))
    quantized_layers = [l for l in all_layers if l not in exclude_layers]

    override_options = {}
    for opt in config.get('override_options', []):
        for regexp in opt['layers']:
            for l in get_layers(model, filter_regexp=regexp):
                assert l not in override_options
                override_options[l] = opt['options']
    for layer in quantized_layers:
        module = attrgetter(layer)(model)
        weight_config, input_config = default_weight_config, default_input_config
        if layer in override_options:
            weight_config = _override_options(weight_config, override_options[layer].get("weight", None))
            input_config = _override_options(input_config, override_options[layer].get("input", None))
        layer_config = {"input": copy.deepcopy(input_config), "weight": copy.deepcopy(weight_config),
                        "momentum": default_momentum_config}
        new_type = mapping[type(module)]
        module_q = new_type.__new__(new_type)
        module_q.__dict__.update(module.__dict__)
        module_q.initialize(layer_config)
        attrsetter(layer)(model, module_q)
        if logger is not None:
            logger.info("Quantize module {} with method <input: {}> <weight: {}>".format(
                layer, layer_config['input'], layer_config['weight']))
