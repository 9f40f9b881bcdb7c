"""RootQLinear (reference: dlmc/quantization/scalar/RootQ/linear.py); the class body lives in scalar/_layers.py."""
from .._layers import make_layer
from .base import RootQBase

__all__ = ["RootQLinear"]
RootQLinear = make_layer("RootQLinear", RootQBase, "linear", __name__)
