from .conv import *  # noqa: F401,F403
from .linear import *  # noqa: F401,F403
from .base import *  # noqa: F401,F403
from .conv import QConv2d  # noqa: F401
from .linear import QLinear  # noqa: F401
from .base import QBase  # noqa: F401
from .group import group_weight_quantizers  # noqa: F401
