from .broken_thruster import BrokenThruster  # noqa: F401
from .controller_debug import ControllerDebug, DebugVal  # noqa: F401
