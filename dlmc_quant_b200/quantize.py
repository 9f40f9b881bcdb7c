"""dlmc/utils/quantize.py::quantize_model - the plug-in boundary: swaps nn.Conv2d / nn.Linear instances
for the quantised classes in place (class swap via __new__ + __dict__.update + initialize, exactly like
quantize.py:130-136 - __init__ is never called), with the same config schema (`weight` / `input` blocks,
`exclude_layers` regexes, `override_options`, `momentum`).  Host-side string plumbing only.

BitMixer / MetaQ mappings are not offered: those packages are absent from the reference tree as well
(quantize.py:10,12 import modules that do not exist)."""
import copy
import re
from operator import attrgetter
from typing import Dict, Iterable, List

from torch import nn

from .scalar import FSPTQuant as FSPQ
from .scalar import RootQ as RQ
from .scalar import modules as qnn
from .scalar.modules.group import group_weight_quantizers  # noqa: F401  (step-level weight grouping)

__all__ = ['quantize_model', 'get_layers', 'attrsetter', 'group_weight_quantizers']

# quantization_type -> {torch layer class: quantised class}   (quantize.py:19-42; None = the QAT / PTQ / LSQ modules)
FAMILIES = {
    None: {nn.Conv2d: qnn.QConv2d, nn.Linear: qnn.QLinear},
    "RootQ": {nn.Conv2d: RQ.RootQConv2d, nn.Linear: RQ.RootQLinear},
    "FSPTQ": {nn.Conv2d: FSPQ.FSPTQConv2d, nn.Linear: FSPQ.FSPTQLinear},
}
MODULE_MAPPING, ROOTQ_MAPPING, FSPTQUANT_MAPPING = FAMILIES[None], FAMILIES["RootQ"], FAMILIES["FSPTQ"]


def attrsetter(*paths):
    """dlmc/utils/access.py:12-27: setter for dotted attribute paths (`attrsetter("layer1.0.conv1")(model, m)`)."""
    def setter(root, value):
        for path in paths:
            *parents, leaf = path.split(".")
            owner = root
            for name in parents:
                owner = getattr(owner, name)
            setattr(owner, leaf, value)
    return setter


def get_layers(model: nn.Module, filter_regexp: str = "(.*?)", filter_types: Iterable = None) -> List[str]:
    """dlmc/utils/access.py:30-61: the names of the modules that own a weight (derived from the parameter names,
    `.weight` / spectral-norm `.weight_orig` stripped, biases skipped), kept when they match the regular expression
    (an optional DataParallel `module.` prefix is tolerated) and, if given, the module types."""
    pattern = re.compile(r"(module\.)?(" + filter_regexp + ")")
    names = []
    for pname, _ in model.named_parameters():
        if "bias" in pname:
            continue
        name = pname.replace(".weight_orig", "").replace(".weight", "")
        if pattern.match(name):
            names.append(name)
    if filter_types is not None:
        names = [n for n in names if isinstance(attrgetter(n)(model), filter_types)]
    return names


def _merged(base: Dict, override: Dict = None) -> Dict:
    """quantize.py:44-58: `type`, `enable` replace, `args` update; the default block is never modified."""
    merged = copy.deepcopy(base)
    for key in ("type", "enable"):
        if override and key in override:
            merged[key] = override[key]
    if override and "args" in override:
        merged["args"].update(override["args"])
    return merged


def _swap_class(model, name, quantised_cls, layer_config):
    """quantize.py:130-136: same object state under the quantised class - __init__ is never called."""
    original = attrgetter(name)(model)
    swapped = quantised_cls.__new__(quantised_cls)
    swapped.__dict__.update(original.__dict__)
    swapped.initialize(layer_config)
    attrsetter(name)(model, swapped)
    return swapped


def quantize_model(model: nn.Module, config: Dict, logger=None, quantization_type: str = None, **kwargs) -> None:
    """quantize.py:61-142: every Conv2d / Linear that is not excluded becomes the quantised class of the chosen
    family, configured from the `weight` / `input` blocks plus the first matching `override_options` entry."""
    if quantization_type in ("BitMixer", "MetaQ"):
        raise NotImplementedError(f"{quantization_type}: its package is not part of the reference tree either")
    family = FAMILIES.get(quantization_type, FAMILIES[None])
    momentum = config['momentum'] if quantization_type == "RootQ" else 0.1

    excluded = {n for rx in config.get('exclude_layers', []) for n in get_layers(model, filter_regexp=rx)}
    overrides = {}
    for entry in config.get('override_options', []):
        for rx in entry['layers']:
            for n in get_layers(model, filter_regexp=rx):
                assert n not in overrides, f"layer {n} matched by two override_options entries"
                overrides[n] = entry['options']

    for name in get_layers(model, filter_types=tuple(family)):
        if name in excluded:
            continue
        opts = overrides.get(name, {})
        layer_config = {"input": _merged(config['input'], opts.get("input")),
                        "weight": _merged(config['weight'], opts.get("weight")), "momentum": momentum}
        _swap_class(model, name, family[type(attrgetter(name)(model))], layer_config)
        if logger is not None:
            logger.info("Quantize module {} with method <input: {}> <weight: {}>".format(
                name, layer_config['input'], layer_config['weight']))
