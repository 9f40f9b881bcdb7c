"""RootQConv2d (reference: dlmc/quantization/scalar/RootQ/conv.py); the class body lives in scalar/_layers.py."""
from .._layers import make_layer
from .base import RootQBase

__all__ = ["RootQConv2d"]
RootQConv2d = make_layer("RootQConv2d", RootQBase, "conv", __name__)
