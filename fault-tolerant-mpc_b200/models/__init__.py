from .sys_model import SystemModel  # noqa: F401
from .spiral_model import SpiralModel  # noqa: F401
