from .conv import *  # noqa: F401,F403
from .linear import *  # noqa: F401,F403
from .base import *  # noqa: F401,F403
from .conv import FSPTQConv2d  # noqa: F401
from .linear import FSPTQLinear  # noqa: F401
from .base import FSPTQBase  # noqa: F401
