from .conv import *  # noqa: F401,F403
from .linear import *  # noqa: F401,F403
from .base import *  # noqa: F401,F403
from .conv import RootQConv2d  # noqa: F401
from .linear import RootQLinear  # noqa: F401
from .base import RootQBase  # noqa: F401
