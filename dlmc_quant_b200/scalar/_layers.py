"""The six quantised layer classes (QConv2d / QLinear, RootQConv2d / RootQLinear, FSPTQConv2d / FSPTQLinear) differ
only in the quantizer base they inherit; what they share is defined once here.

Reference: dlmc/quantization/scalar/{modules,RootQ,FSPTQuant}/{conv,linear}.py - each is `class X(Base, nn.Conv2d)`
whose `_forward_func(input, weight)` is the plain library call (the observers call it with a quantised weight,
ops.py:86,100,254,272).  The convolution / matmul itself stays a cuDNN / cuBLAS call: only the fake-quant around it
is on this repo's hot path."""
import torch.nn.functional as F
from torch import nn
from torch.nn.modules.utils import _pair


def _conv_forward(self, input, weight):
    if self.padding_mode == 'zeros':
        return F.conv2d(input, weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
    padded = F.pad(input, self._reversed_padding_repeated_twice, mode=self.padding_mode)
    return F.conv2d(padded, weight, self.bias, self.stride, _pair(0), self.dilation, self.groups)


def _linear_forward(self, input, weight):
    return F.linear(input, weight, self.bias)


_KINDS = {"conv": (nn.Conv2d, _conv_forward), "linear": (nn.Linear, _linear_forward)}


def make_layer(name, quantizer_base, kind, module):
    """`class name(quantizer_base, nn.Conv2d | nn.Linear)` with the reference's constructor convention
    (`X(*torch_args, qconfig=...)`; `quantize_model` never calls it - it swaps classes via __new__ + initialize)."""
    torch_cls, forward = _KINDS[kind]

    def __init__(self, *args, qconfig=None, **kwargs):
        torch_cls.__init__(self, *args, **kwargs)
        self.initialize(qconfig)

    return type(name, (quantizer_base, torch_cls),
                {"__init__": __init__, "_forward_func": forward, "__module__": module,
                 "__doc__": f"{quantizer_base.__name__} quantizers around torch.nn.{torch_cls.__name__}."})
