"""dlmc/quantization/scalar/modules/linear.py: QLinear."""
import torch.nn.functional as F
from torch.nn import Linear

from .base import QBase


class QLinear(QBase, Linear):
    def __init__(self, *args, qconfig=None, **kwargs):
        Linear.__init__(self, *args, **kwargs)
        self.initialize(qconfig)

    def _forward_func(self, input, weight):
        return F.linear(input, weight, self.bias)
