"""QConv2d (reference: dlmc/quantization/scalar/modules/conv.py); the class body lives in scalar/_layers.py."""
from .._layers import make_layer
from .base import QBase

__all__ = ["QConv2d"]
QConv2d = make_layer("QConv2d", QBase, "conv", __name__)
