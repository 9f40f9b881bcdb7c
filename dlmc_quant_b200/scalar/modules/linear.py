"""QLinear (reference: dlmc/quantization/scalar/modules/linear.py); the class body lives in scalar/_layers.py."""
from .._layers import make_layer
from .base import QBase

__all__ = ["QLinear"]
QLinear = make_layer("QLinear", QBase, "linear", __name__)
