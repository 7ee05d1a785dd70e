"""SpiralModel -- host-side mirror of ft_mpc/models/spiral_model.py (orbit-centre prediction model).

State order [p_c(3) v_c(3) w(3) q(4)]; only the first 9 states are controlled.  The dynamics, RK4 and
their derivatives live in csrc/ftmpc_dyn.cuh; this class only carries the parameters.
"""
import numpy as np

from .sys_model import SystemModel


class SpiralModel(SystemModel):
    def __init__(self, dt, spiral_params):
        self.r = spiral_params.r
        self.spiral_params = spiral_params
        super().__init__(dt)

    @classmethod
    def from_system_model(cls, sys_model):
        """spiral_model.py:31-42"""
        from ..controllers.tools.spiral_parameters import SpiralParameters
        new = cls(sys_model.dt, SpiralParameters(sys_model))
        for bt in sys_model.broken_thrusters:
            new.set_fault(bt)
        return new

    def normalize_quaternion(self, state):
        state = np.array(state, dtype=float).reshape(-1)
        state[9:13] = state[9:13] / np.linalg.norm(state[9:13])
        return state

    def robot_to_center(self, x):
        """Robot state -> centre state (spiral_model.py:91-109); host-side convenience in numpy.
        The per-step path does this transform inside ftmpc_step (csrc/ftmpc_dyn.cuh: robot_to_center)."""
        x = np.asarray(x, float).reshape(-1)
        q, w = x[6:10], x[10:13]
        qx, qy, qz, qw = q
        R = np.array([[qx*qx - qy*qy - qz*qz + qw*qw, 2*(qx*qy + qz*qw), 2*(qx*qz - qy*qw)],
                      [2*(qx*qy - qz*qw), -qx*qx + qy*qy - qz*qz + qw*qw, 2*(qy*qz + qx*qw)],
                      [2*(qx*qz + qy*qw), 2*(qy*qz - qx*qw), -qx*qx - qy*qy + qz*qz + qw*qw]])
        return np.concatenate((x[0:3] + R.T @ self.r, x[3:6] + R.T @ np.cross(w, self.r), w, q))

    @property
    def Nu(self):
        return self.Nu_simplified
