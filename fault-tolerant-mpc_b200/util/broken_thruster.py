"""Fault record -- mirrors ft_mpc/util/broken_thruster.py:1-10 of the reference."""


class BrokenThruster:
    """A failed thruster: `index` in 0..15, `intensity` in [0,1] (0 = dead, 1 = stuck fully open;
    intensity*max_thrust is the stuck-on force, sys_model.py:239)."""

    def __init__(self, index, intensity):
        self.index = int(index)
        self.intensity = float(intensity)

    def __repr__(self):
        return f"BrokenThruster(index={self.index}, intensity={self.intensity})"
