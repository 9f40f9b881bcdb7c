"""FSPTQConv2d (reference: dlmc/quantization/scalar/FSPTQuant/conv.py); the class body lives in scalar/_layers.py."""
from .._layers import make_layer
from .base import FSPTQBase

__all__ = ["FSPTQConv2d"]
FSPTQConv2d = make_layer("FSPTQConv2d", FSPTQBase, "conv", __name__)
