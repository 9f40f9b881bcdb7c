"""FSPTQLinear (reference: dlmc/quantization/scalar/FSPTQuant/linear.py); the class body lives in scalar/_layers.py."""
from .._layers import make_layer
from .base import FSPTQBase

__all__ = ["FSPTQLinear"]
FSPTQLinear = make_layer("FSPTQLinear", FSPTQBase, "linear", __name__)
