"""TEST INFRASTRUCTURE ONLY - import shim for the unmodified reference.

Loads the hot-path modules of ilur98/DLMC-QUANT straight from `/root/reference`
(read-only, only present in the authoring container - never on the GPU box) so
that `oracle/make_golden.py` can run the reference's *own* code to mint golden
vectors and `tests/test_oracle_vs_reference.py` can pin `oracle/restate.py`
against it.  Nothing in the product package may import this file.

Why a shim is needed:
  * dlmc/quantization/scalar/ops.py:2-3 and FSPTQuant/base.py:5 import
    matplotlib (unused) -> stubbed with empty modules carrying the two names.
  * ops.py:5 imports trainer.loss.loss.l2_loss, but trainer/__init__.py:2-7
    imports trainer modules that are not in the repo -> `trainer` and
    `trainer.loss` are pre-seeded as namespace stubs whose __path__ points at
    the real directories, so trainer/loss/loss.py itself loads unmodified.
"""
import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DLMCQ_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dlmc", "quantization", "scalar"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


_loaded = None


def load():
    """Return a namespace with the reference's hot-path modules:
    .utils .ops .modules .RootQ .FSPTQuant .loss  (all unmodified reference code)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "matplotlib" not in sys.modules:
        m = _stub("matplotlib", scale=None)
        m.__path__ = []
        _stub("matplotlib.pyplot", sca=None)
    if "trainer" not in sys.modules:
        t = _stub("trainer")
        t.__path__ = [os.path.join(REFERENCE_ROOT, "trainer")]
        tl = _stub("trainer.loss")
        tl.__path__ = [os.path.join(REFERENCE_ROOT, "trainer", "loss")]
    ns = types.SimpleNamespace()
    ns.loss = importlib.import_module("trainer.loss.loss")
    ns.utils = importlib.import_module("dlmc.quantization.scalar.utils")
    ns.ops = importlib.import_module("dlmc.quantization.scalar.ops")
    ns.modules = importlib.import_module("dlmc.quantization.scalar.modules")
    ns.RootQ = importlib.import_module("dlmc.quantization.scalar.RootQ")
    ns.rootq_function = importlib.import_module("dlmc.quantization.scalar.RootQ.function")
    ns.modules_function = importlib.import_module("dlmc.quantization.scalar.modules.function")
    ns.FSPTQuant = importlib.import_module("dlmc.quantization.scalar.FSPTQuant")
    _loaded = ns
    return ns


_reparam = None


def load_reparam():
    """The reference's weight-space re-parameterisation code, unmodified:
    .merge_bn  = dlmc/utils/merge_bn.py   (its `dlmc.utils` package __init__ pulls in timm / BitMixer /
                 MetaQuant, none of which exist here -> `dlmc.utils` is pre-seeded as a namespace stub and
                 `dlmc.quantization.scalar.BitMixer` as a stub carrying the two imported names)
    .repvgg    = model/classification/repvgg.py (imports only torch / numpy; loaded by path because
                 model/classification/__init__.py imports timm)."""
    global _reparam
    if _reparam is not None:
        return _reparam
    load()
    if "dlmc.utils" not in sys.modules:
        u = _stub("dlmc.utils")
        u.__path__ = [os.path.join(REFERENCE_ROOT, "dlmc", "utils")]
    if "dlmc.quantization.scalar.BitMixer" not in sys.modules:
        _stub("dlmc.quantization.scalar.BitMixer", BitMixerBatchNorm=None, BitMixerSwitchableBatchNorm=None)
    ns = types.SimpleNamespace()
    ns.merge_bn = importlib.import_module("dlmc.utils.merge_bn")
    spec = importlib.util.spec_from_file_location(
        "_ref_repvgg", os.path.join(REFERENCE_ROOT, "model", "classification", "repvgg.py"))
    ns.repvgg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ns.repvgg)
    _reparam = ns
    return ns
