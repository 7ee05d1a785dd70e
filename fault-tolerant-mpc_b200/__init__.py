"""ft_mpc_b200 -- B200-native drop-in for the per-timestep MPC solve of DISCOWER/fault-tolerant-mpc.

The directory name carries a hyphen (``fault-tolerant-mpc_b200``), so it is imported under the name
``ft_mpc_b200`` through ``ftmpc_import.py`` at the repository root::

    import ftmpc_import; ft = ftmpc_import.load()
    from ft_mpc_b200.models import SystemModel, SpiralModel
    from ft_mpc_b200.controllers import SpiralingController

Everything numeric on the per-step path runs in ``csrc/libftmpc.so`` (CUDA, sm_100a) behind the C ABI of
``include/ftmpc.h``.  There is no CPU fallback: constructing a controller without the library or
without a CUDA device raises.
"""
__all__ = ["models", "controllers", "util", "distributed", "_lib"]
