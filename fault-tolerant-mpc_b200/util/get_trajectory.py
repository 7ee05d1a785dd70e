"""Reference-trajectory generator (host-side input preparation).

Restates the parts of ft_mpc/util/get_trajectory.py:71-184 the demo scenario needs: `hover`
(point stabilising, optional `hover_<x>_<y>_<z>`), `generate_line` and `generate_circle`.
Returns a 13 x T robot-state reference [p v q w] sampled at dt over 10*duration seconds.
Signature as in the reference: load_trajectory(action, dt, duration=100, file_path=None)  (get_trajectory.py:6).
"""
import numpy as np


def load_trajectory(action: str, dt: float, duration: float = 100, file_path=None) -> np.ndarray:
    if action == "load":
        raise NotImplementedError("trajectory files ('load') are host-side input preparation outside the hot path (SURVEY.md section 2 row 9)")
    t = np.arange(0, 10 * duration, dt).reshape(1, -1)
    one, zero = np.ones(t.shape), np.zeros(t.shape)
    ident_q = np.vstack((zero, zero, zero, one))                  # identity quaternion [x y z w]
    rates = np.zeros((3, t.size))
    if action in ("hover", "generate_point_stabilizing") or action.startswith("hover_"):
        pos = [0.0, 0.0, 0.0]
        if action.startswith("hover_"):
            params = action.split("_")[1:]
            if len(params) != 3:
                raise ValueError(f"Invalid number of parameters for action '{action}'")
            pos = [float(p) for p in params]
        return np.concatenate((pos[0] * one, pos[1] * one, pos[2] * one, zero, zero, zero, ident_q, rates))
    if action == "generate_line":
        return np.concatenate((t, zero, zero, one, zero, zero, ident_q, rates))
    if action == "generate_circle" or action.startswith("circle_"):
        radius, s_per_rot = 2.0, 30.0
        if action.startswith("circle_"):
            p = action.split("_")[1:]
            if len(p) != 4 or p[0] != "r" or p[2] != "sPerFullCircle":
                raise ValueError(f"Invalid parameters for action '{action}'")
            radius, s_per_rot = float(p[1]), float(p[3])
        om = 2 * np.pi / s_per_rot
        x = np.concatenate((radius * np.cos(om * t), radius * np.sin(om * t), zero,
                            -radius * om * np.sin(om * t), radius * om * np.cos(om * t), zero, ident_q, rates))
        x += np.array([-radius] + [0] * 12).reshape(-1, 1)
        return x
    raise ValueError(f"Invalid action '{action}'.")
