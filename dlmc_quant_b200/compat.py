"""Import-path compatibility: make `import dlmc.quantization.scalar...` / `import dlmc.utils.quantize` resolve to this
package, so that code written against the reference's module paths - its trainers
(`trainer/quantization_aware_training_trainer.py`, `trainer/fsptq_trainer.py`) and example scripts - picks up the
B200 modules without an edit:

    import dlmc_quant_b200.compat as compat
    compat.install()                     # before the first `import dlmc...`
    from dlmc.utils.quantize import quantize_model
    from dlmc.quantization.scalar import modules as qnn, RootQ, FSPTQuant, ops, utils

Only the hot-path modules are aliased (the reference paths on the left):
    dlmc.quantization.scalar{,.utils,.ops,.modules,.RootQ,.FSPTQuant}   -> dlmc_quant_b200.scalar...
    dlmc.utils.quantize, dlmc.utils.merge_bn                            -> dlmc_quant_b200.quantize / .reparam
Everything else under `dlmc` keeps resolving to whatever is on sys.path (the reference checkout, if present)."""
import importlib
import sys
import types

_ALIASES = {
    "dlmc.quantization.scalar": "dlmc_quant_b200.scalar",
    "dlmc.quantization.scalar.utils": "dlmc_quant_b200.scalar.utils",
    "dlmc.quantization.scalar.ops": "dlmc_quant_b200.scalar.ops",
    "dlmc.quantization.scalar.modules": "dlmc_quant_b200.scalar.modules",
    "dlmc.quantization.scalar.RootQ": "dlmc_quant_b200.scalar.RootQ",
    "dlmc.quantization.scalar.FSPTQuant": "dlmc_quant_b200.scalar.FSPTQuant",
    # submodules the reference imports by full path (e.g. trainer/fsptq_trainer.py:9 `from dlmc.quantization.scalar.
    # FSPTQuant.base import FSPTQBase`): they must be the SAME module objects, or isinstance() checks see two classes
    "dlmc.quantization.scalar.modules.base": "dlmc_quant_b200.scalar.modules.base",
    "dlmc.quantization.scalar.modules.function": "dlmc_quant_b200.scalar.modules.function",
    "dlmc.quantization.scalar.modules.conv": "dlmc_quant_b200.scalar.modules.conv",
    "dlmc.quantization.scalar.modules.linear": "dlmc_quant_b200.scalar.modules.linear",
    "dlmc.quantization.scalar.RootQ.base": "dlmc_quant_b200.scalar.RootQ.base",
    "dlmc.quantization.scalar.RootQ.function": "dlmc_quant_b200.scalar.RootQ.function",
    "dlmc.quantization.scalar.RootQ.conv": "dlmc_quant_b200.scalar.RootQ.conv",
    "dlmc.quantization.scalar.RootQ.linear": "dlmc_quant_b200.scalar.RootQ.linear",
    "dlmc.quantization.scalar.FSPTQuant.base": "dlmc_quant_b200.scalar.FSPTQuant.base",
    "dlmc.quantization.scalar.FSPTQuant.conv": "dlmc_quant_b200.scalar.FSPTQuant.conv",
    "dlmc.quantization.scalar.FSPTQuant.linear": "dlmc_quant_b200.scalar.FSPTQuant.linear",
    "dlmc.utils.quantize": "dlmc_quant_b200.quantize",
    "dlmc.utils.merge_bn": "dlmc_quant_b200.reparam",
}


def _namespace(name):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []          # a package, so that `import name.sub` is attempted through sys.modules
        sys.modules[name] = mod
    return mod


def install(force=False):
    """Register the aliases.  Existing entries (an already imported reference) are kept unless force=True."""
    for alias, target in _ALIASES.items():
        if alias in sys.modules and not force:
            continue
        parts = alias.split(".")
        for k in range(1, len(parts)):
            _namespace(".".join(parts[:k]))
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    return sorted(_ALIASES)


def uninstall():
    for alias in _ALIASES:
        if sys.modules.get(alias) is not None and sys.modules[alias].__name__.startswith("dlmc_quant_b200"):
            del sys.modules[alias]
