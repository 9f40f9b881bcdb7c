"""Host-side mirror of dlmc/quantization/scalar: same module / function names, CUDA kernels inside."""
from . import ops, utils  # noqa: F401
