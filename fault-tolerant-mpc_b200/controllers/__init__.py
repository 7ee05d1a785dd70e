from .spiraling_mpc import SpiralingController  # noqa: F401
