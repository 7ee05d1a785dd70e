from .broken_thruster import BrokenThruster  # noqa: F401
