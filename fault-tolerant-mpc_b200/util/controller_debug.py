"""Per-step debug records and the result export format of the reference.

Mirrors ft_mpc/util/controller_debug.py: `DebugVal` (:9-79, field names and error definitions) and
`ControllerDebug.add_debug_val / get_time / export` (:81-90, :216-260).  The export is the reference's
67-column ';'-separated CSV (numpy.savetxt with a '# '-prefixed header line), so the reference's plotting /
animation tools read files written here.  The matplotlib views (show_* / animate_3d) are out of scope.

`ControllerDebug.from_batch` builds the same history for ONE instance of a batched closed-loop rollout
from recorded tensors, so a Monte-Carlo run can be exported without going through get_control.
"""
import copy

import numpy as np

HEADER = (["time"]
          + [f"position_{a}" for a in "xyz"] + [f"velocity_{a}" for a in "xyz"]
          + [f"orientation_{a}" for a in "xyzw"] + [f"angular_velocity_{a}" for a in "xyz"]
          + [f"input_{i}" for i in range(16)]
          + [f"force_{a}" for a in "xyz"] + [f"torque_{a}" for a in "xyz"]
          + [f"circle_position_{a}" for a in "xyz"] + [f"circle_velocity_{a}" for a in "xyz"]
          + [f"circle_angular_velocity_{a}" for a in "xyz"]
          + [f"position_error_{a}" for a in "xyz"] + [f"velocity_error_{a}" for a in "xyz"]
          + [f"orientation_error_{a}" for a in "xyzw"] + [f"angular_velocity_error_{a}" for a in "xyz"]
          + [f"circle_position_error_{a}" for a in "xyz"] + [f"circle_velocity_error_{a}" for a in "xyz"]
          + [f"circle_angular_velocity_error_{a}" for a in "xyz"])          # controller_debug.py:240-257

_ROW_FIELDS = ("position", "velocity", "orientation", "angular_velocity", "input", "force", "torque",
               "circle_position", "circle_velocity", "circle_angular_velocity",
               "position_error", "velocity_error", "orientation_error", "angular_velocity_error",
               "circle_position_error", "circle_velocity_error", "circle_angular_velocity_error")   # :219-226


class DebugVal:
    """One closed-loop step (controller_debug.py:9-79)."""

    def __init__(self, controller, t):
        self.controller = str(controller)
        model = getattr(controller, "model", None)
        self.faulty_force = copy.deepcopy(getattr(model, "faulty_force", None))
        self.time = t
        for f in _ROW_FIELDS:
            setattr(self, f, None)
        self.desired_position = self.desired_velocity = None
        self.desired_orientation = self.desired_angular_velocity = None

    def set_state(self, x):                                                   # robot state [p v q w]     :39-44
        x = np.array(x, dtype=float).flatten()
        self.position, self.velocity, self.orientation, self.angular_velocity = x[0:3], x[3:6], x[6:10], x[10:13]

    def set_circle_state(self, c):                                            # centre state [p_c v_c w]  :46-50
        c = np.array(c, dtype=float).flatten()
        self.circle_position, self.circle_velocity, self.circle_angular_velocity = c[0:3], c[3:6], c[6:9]

    def set_input(self, u, model):                                            # :52-57
        u = np.array(u, dtype=float).flatten()
        self.input = u
        g = np.asarray(model.D, dtype=float) @ u
        self.force, self.torque = g[0:3], g[3:6]

    def set_desired_state(self, x):                                           # :59-67
        x = np.array(x, dtype=float).flatten()
        self.desired_position, self.desired_velocity = x[0:3], x[3:6]
        if x.size == 9:
            self.desired_angular_velocity = x[6:9]
            self.desired_orientation = np.zeros(4)
        else:
            self.desired_orientation = x[6:10]
            self.desired_angular_velocity = x[10:13]

    def calculate_errors(self):                                               # :69-79
        if self.position is not None and self.desired_position is not None:
            self.position_error = self.desired_position - self.position
            self.velocity_error = self.desired_velocity - self.velocity
            self.orientation_error = self.desired_orientation - self.orientation
            self.angular_velocity_error = self.desired_angular_velocity - self.angular_velocity
        if self.circle_position is not None and self.desired_position is not None:
            self.circle_position_error = self.desired_position - self.circle_position
            self.circle_velocity_error = self.desired_velocity - self.circle_velocity
            self.circle_angular_velocity_error = self.desired_angular_velocity - self.circle_angular_velocity

    def row(self) -> np.ndarray:
        return np.concatenate([np.array([self.time], dtype=float)] + [np.asarray(getattr(self, f), float).flatten()
                                                                      for f in _ROW_FIELDS])


class ControllerDebug:
    def __init__(self):
        self.history = []

    def add_debug_val(self, debug_val):
        self.history.append(debug_val)

    def get_time(self):
        return [h.time for h in self.history]

    def table(self) -> np.ndarray:
        """[len(history), 67] array in the column order of HEADER"""
        return np.stack([h.row() for h in self.history]) if self.history else np.zeros((0, len(HEADER)))

    def export(self, file_path):
        """`file_path + ".csv"`, ';'-separated, header as written by the reference (controller_debug.py:216-260)."""
        np.savetxt(str(file_path) + ".csv", self.table(), delimiter=";", header=";".join(HEADER))

    @classmethod
    def from_batch(cls, controller, times, states, centers, thrusts, desired):
        """History of one instance from recorded arrays: times [T], states [T,13] robot states, centers [T,>=9]
        centre states, thrusts [T,16], desired [T,9] reference points."""
        dbg = cls()
        for k in range(len(times)):
            dv = DebugVal(controller, float(times[k]))
            dv.set_state(states[k])
            dv.set_circle_state(centers[k])
            dv.set_input(thrusts[k], controller.model)
            dv.set_desired_state(desired[k])
            dv.calculate_errors()
            dbg.add_debug_val(dv)
        return dbg
