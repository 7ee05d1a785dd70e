"""SystemModel -- host-side mirror of ft_mpc/models/sys_model.py (constants, D, fault bookkeeping).

The numerics (RK4 of dx_dt) run on the GPU: `dynamics(x, u)` calls ftmpc_plant_step (csrc/ftmpc_plant.cuh).
"""
import numpy as np


def allocation_matrix() -> np.ndarray:
    """D (6x16), rows [Fx Fy Fz tx ty tz].  sys_model.py:73-123"""
    D = np.zeros((6, 16))
    d1, d2, d3 = 0.12, 0.09, 0.05
    D[0, [0, 1, 4, 5]], D[0, [2, 3, 6, 7]] = -1.0, 1.0
    D[1, [8, 9]], D[1, [10, 11]] = -1.0, 1.0
    D[2, [12, 14]], D[2, [13, 15]] = -1.0, 1.0
    D[3, [12, 15]], D[3, [13, 14]] = -d1, d1
    D[4, [0, 3, 4, 7]], D[4, [1, 2, 5, 6]] = -d3, d3
    D[5, [0, 1, 6, 7]], D[5, [2, 3, 4, 5]] = d1, -d1
    D[5, [8, 11]], D[5, [9, 10]] = -d2, d2
    return D


class SystemModel:
    """3-D rigid body with 16 thrusters; robot state [p(3) v(3) q(4) w(3)], q = [x y z w]."""

    def __init__(self, dt):
        self.mass = 16.8                                   # sys_model.py:52
        self.inertia = np.diag([0.2, 0.3, 0.25])           # :53-57
        self.inertia_inv = np.linalg.inv(self.inertia)
        self.max_thrust = 3.4                              # :60
        self.Nx, self.Nu_simplified, self.Nu_full = 13, 6, 16
        self.dt = dt
        self.D = allocation_matrix()
        self.broken_thrusters = []
        self.faulty_force = np.zeros(self.Nu_full)
        self.faulty_force_generalized = self.D @ self.faulty_force
        self.u_ub_physical = np.array([self.max_thrust] * self.Nu_full)
        self._plant = None

    def set_fault(self, broken_thruster):
        """sys_model.py:228-243"""
        self.broken_thrusters.append(broken_thruster)
        self.faulty_force = np.zeros(self.Nu_full)
        self.u_ub_physical = np.array([self.max_thrust] * self.Nu_full)
        for th in self.broken_thrusters:
            self.faulty_force[th.index] = th.intensity * self.max_thrust
            self.u_ub_physical[th.index] = 0.0
        self.faulty_force_generalized = self.D @ self.faulty_force

    @property
    def fault_mask(self) -> int:
        m = 0
        for th in self.broken_thrusters:
            m |= 1 << th.index
        return m

    @property
    def Nu(self):
        return self.Nu_full

    def normalize_quaternion(self, state):
        """sys_model.py:164-175"""
        state = np.array(state, dtype=float).reshape(-1)
        state[6:10] = state[6:10] / np.linalg.norm(state[6:10])
        return state

    def dynamics(self, x, u):
        """One RK4 step of the 16-input plant on the GPU (sys_model.py:138-226); returns ndarray[13]."""
        from ..controllers.spiraling_mpc import _PlantStepper
        if self._plant is None:
            self._plant = _PlantStepper(self)
        return self._plant(np.asarray(x, float).reshape(1, 13), np.asarray(u, float).reshape(1, 16))[0]
