"""dlmc_quant_b200 - B200-native fake-quantisation kernels behind DLMC-QUANT's quantizer API.

    from dlmc_quant_b200 import quantize_model               # dlmc.utils.quantize.quantize_model
    from dlmc_quant_b200.scalar import modules, RootQ, FSPTQuant, ops, utils

Everything that computes goes through libdlmcq.so (C ABI in include/dlmcq.h, hand-written sm_100a
CUDA).  There is no CPU fallback: without the library or without a CUDA device the ops raise."""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "quantize_model":
        from .quantize import quantize_model
        return quantize_model
    if name in ("functional", "scalar", "dist", "quantize", "build", "reparam", "recon", "fuse", "graph", "calibrate"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
