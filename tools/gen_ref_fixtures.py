#!/usr/bin/env python
"""Generate tests/golden/ref_fixtures.npz by EXECUTING THE REFERENCE'S OWN PYTHON MODULES in the build container.

The reference (read-only at /root/reference) needs casadi / cvxpy / polytope / matplotlib, none of which is installable
here.  Its model code only uses a small, purely algebraic part of the CasADi API (MX.sym / MX.zeros / slicing /
vertcat / cross / DM / Function), so this script registers a ~100-line numeric stand-in for `casadi` (lazy expression
objects evaluated with numpy) and empty stubs for the plotting / optimisation packages that are imported but not
reached, then imports

    ft_mpc.util.utils            Rot, RotInv, RotFull, RotFullInv
    ft_mpc.util.broken_thruster  BrokenThruster
    ft_mpc.models.sys_model      SystemModel  (D, set_fault, RK4 plant `dynamics`, normalize_quaternion)
    ft_mpc.models.spiral_model   SpiralModel  (from_system_model -> SpiralParameters -> InputBounds/Qhull, RK4 `dynamics`,
                                               robot_to_center)
    ft_mpc.util.get_trajectory   load_trajectory

and records their outputs on seeded inputs.  Nothing of the reference is copied: only numbers leave this script.
Not executed: the NLP solve (IPOPT), the allocator (CVXPY/OSQP), `load_terminal_ingredients` (it `eval`s the YAML; the
stored expression is parsed with sympy in tools/gen_terminal_data.py instead).
Usage: python tools/gen_ref_fixtures.py        (needs /root/reference; run in the build container only)
"""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


# ---------------------------------------------------------------------------------------------------------
# numeric stand-in for the subset of casadi the reference's model code touches
# ---------------------------------------------------------------------------------------------------------
def _val(x, env):
    if isinstance(x, MX):
        return x.ev(env)
    a = np.asarray(x, dtype=float)
    return a.reshape(-1, 1) if a.ndim <= 1 else a


class MX:
    """lazy matrix expression: `ev(env)` returns a 2-D float array"""
    __array_ufunc__ = None          # let `ndarray @ MX`, `ndarray + MX` dispatch to the reflected operators below

    def __init__(self, fn, shape):
        self.fn, self.shape, self.assign = fn, shape, []

    def ev(self, env):
        key = ("#", id(self))                # shared sub-expressions (RK4 stages) are evaluated once per call
        if key in env:
            return env[key]
        a = np.array(self.fn(env), dtype=float)
        for idx, v in self.assign:
            a[idx] = np.squeeze(_val(v, env)) if np.ndim(a[idx]) == 0 else _val(v, env).reshape(np.shape(a[idx]))
        env[key] = a
        return a

    @staticmethod
    def sym(name, r, c=1):
        return MX(lambda env: env[name].reshape(r, c), (r, c))

    @staticmethod
    def zeros(r, c=1):
        return MX(lambda env: np.zeros((r, c)), (r, c))

    def size1(self):
        return self.shape[0]

    @property
    def T(self):
        return MX(lambda env: self.ev(env).T, self.shape[::-1])

    def __getitem__(self, idx):
        def f(env):
            a = self.ev(env)
            if not isinstance(idx, tuple) and a.shape[1] == 1:
                r = a[:, 0][idx]
            else:
                r = a[idx]
            r = np.asarray(r, dtype=float)
            return r.reshape(-1, 1) if r.ndim <= 1 else r
        probe = np.zeros(self.shape)
        r = probe[:, 0][idx] if (not isinstance(idx, tuple) and self.shape[1] == 1) else probe[idx]
        r = np.asarray(r)
        return MX(f, (r.size, 1) if r.ndim <= 1 else r.shape)

    def __setitem__(self, idx, v):
        self.assign.append((idx, v))

    def _bin(self, o, op, refl=False):
        def f(env):
            a, b = self.ev(env), _val(o, env)
            return op(b, a) if refl else op(a, b)
        return MX(f, self.shape)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __neg__(self): return MX(lambda env: -self.ev(env), self.shape)
    def __pow__(self, p): return MX(lambda env: self.ev(env) ** p, self.shape)

    def __matmul__(self, o):
        osh = o.shape if isinstance(o, MX) else np.asarray(o).reshape(-1, 1).shape if np.ndim(o) <= 1 else np.shape(o)
        return MX(lambda env: self.ev(env) @ _val(o, env), (self.shape[0], osh[1]))

    def __rmatmul__(self, o):
        return MX(lambda env: _val(o, env) @ self.ev(env), (np.shape(o)[0], self.shape[1]))


class _Function:
    def __init__(self, name, ins, outs, opts=None):
        self.ins, self.outs = ins, outs

    def __call__(self, *args):
        env = {}
        for s, a in zip(self.ins, args):
            s.fn(_Probe(env, np.asarray(a, dtype=float)))
        r = self.outs[0].ev(env)
        return r


class _Probe(dict):
    """captures the symbol name when a `sym` lambda looks itself up"""
    def __init__(self, env, value):
        super().__init__()
        self.env, self.value = env, value

    def __getitem__(self, name):
        self.env[name] = self.value
        return self.value

    def __contains__(self, key):
        return False

    def __setitem__(self, key, v):
        pass


def _install_stubs():
    ca = types.ModuleType("casadi")
    ca.MX = MX
    ca.DM = lambda a: np.asarray(a, dtype=float)
    ca.Function = _Function

    def vertcat(*xs):
        n = sum((x.shape[0] if isinstance(x, MX) else np.size(x)) for x in xs)
        return MX(lambda env: np.vstack([_val(x, env).reshape(-1, 1) for x in xs]), (n, 1))

    def cross(a, b):
        return MX(lambda env: np.cross(_val(a, env).ravel(), _val(b, env).ravel()).reshape(-1, 1), (3, 1))

    ca.vertcat, ca.cross = vertcat, cross
    ca.norm_2 = lambda a: (MX(lambda env: np.array([[np.linalg.norm(_val(a, env))]]), (1, 1)) if isinstance(a, MX)
                           else float(np.linalg.norm(np.asarray(a, dtype=float))))
    sys.modules["casadi"] = ca

    class _Any(types.ModuleType):
        """imported-but-not-reached packages: any attribute is an empty class"""
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {})

    for name in ("casadi.tools", "cvxpy", "polytope", "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "qpsolvers", "pympc", "pympc.geometry",
                 "pympc.geometry.polyhedron", "pympc.dynamics", "pympc.dynamics.discrete_time_systems", "pympc.control",
                 "pympc.control.controllers", "pympc.plot", "control"):
        sys.modules.setdefault(name, _Any(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def main():
    if not REF.exists():
        raise SystemExit("/root/reference is not available: fixtures can only be generated in the build container")
    _install_stubs()
    sys.path.insert(0, str(REF))
    from ft_mpc.models.spiral_model import SpiralModel
    from ft_mpc.models.sys_model import SystemModel
    from ft_mpc.util.broken_thruster import BrokenThruster
    from ft_mpc.util.get_trajectory import load_trajectory
    from ft_mpc.util.utils import Rot, RotFull, RotFullInv, RotInv

    out = {}
    rng = np.random.default_rng(2024)
    from scipy.spatial.transform import Rotation
    K = 12
    q = Rotation.random(K, random_state=7).as_quat()
    out["quat"] = q
    out["Rot"] = np.stack([Rot(qq) for qq in q])
    out["RotInv"] = np.stack([RotInv(qq) for qq in q])
    out["RotFull"] = np.stack([RotFull(qq) for qq in q])
    out["RotFullInv"] = np.stack([RotFullInv(qq) for qq in q])

    dt = 0.1
    for tag, faults in (("default", [(10, 1.0), (11, 1.0)]), ("single3dead", [(3, 0.0)]), ("nofault", [])):
        model = SystemModel(dt)
        for i, a in faults:
            model.set_fault(BrokenThruster(i, a))
        out[f"{tag}::D"] = np.asarray(model.D, float)
        out[f"{tag}::consts"] = np.array([model.mass, model.max_thrust, *np.diag(model.inertia)])
        out[f"{tag}::faulty_force"] = np.asarray(model.faulty_force, float).ravel()
        out[f"{tag}::faulty_force_generalized"] = np.asarray(model.faulty_force_generalized, float).ravel()
        out[f"{tag}::u_ub_physical"] = np.asarray(model.u_ub_physical, float).ravel()
        x = np.concatenate([rng.uniform(-1, 1, (K, 3)), rng.uniform(-.5, .5, (K, 3)), q, rng.uniform(-.5, .5, (K, 3)) + [0, 0, .6]], axis=1)
        u16 = rng.uniform(0, model.max_thrust, (K, 16))
        out[f"{tag}::plant_x"] = x
        out[f"{tag}::plant_u"] = u16
        out[f"{tag}::plant_next"] = np.stack([np.asarray(model.dynamics(x[k], u16[k]), float).ravel() for k in range(K)])   # sim_env.py:85
        out[f"{tag}::plant_next_normalized"] = np.stack([np.asarray(model.normalize_quaternion(np.array(out[f"{tag}::plant_next"][k])), float).ravel()
                                                         for k in range(K)])                                                  # sim_env.py:93
        if tag == "nofault":
            continue                                                     # InputBounds of the healthy vehicle: 2^16 corners, skipped
        spiral = SpiralModel.from_system_model(model)                    # sim.py:33
        sp = spiral.spiral_params
        out[f"{tag}::r"] = np.asarray(spiral.r, float).ravel()
        out[f"{tag}::omega_des"] = np.asarray(sp.omega_des, float).ravel()
        out[f"{tag}::f_virt"] = np.asarray(sp.f_virt, float).ravel()
        out[f"{tag}::compensation_force"] = np.asarray(sp.compensation_force, float).ravel()
        hull = sp.input_bounds.get_conv_hull() if hasattr(sp, "input_bounds") else None
        if hull is None:
            from ft_mpc.controllers.tools.input_bounds import InputBounds
            hull = InputBounds(model).get_conv_hull()
        A, b = (hull.A, hull.b) if hasattr(hull, "A") else hull
        out[f"{tag}::hull_A"], out[f"{tag}::hull_b"] = np.asarray(A, float), np.asarray(b, float).ravel()
        c = np.stack([np.asarray(spiral.robot_to_center(x[k]), float).ravel() for k in range(K)])          # spiral_model.py:91-109
        u6 = rng.normal(0, 2.0, (K, 6))
        out[f"{tag}::center"] = c
        out[f"{tag}::spiral_u"] = u6
        out[f"{tag}::spiral_next"] = np.stack([np.asarray(spiral.dynamics(c[k], u6[k]), float).ravel() for k in range(K)])   # RK4 of spiral_model.py:44-76
    for cmd in ("hover", "hover_1_-2_0.5", "generate_line", "generate_circle"):
        out[f"traj::{cmd}"] = np.asarray(load_trajectory(cmd, dt, 3), float)          # (action, dt, duration), get_trajectory.py:6
    # ---- reference-window code of the controller (spiraling_mpc.py:255-286, 356-365), run UNBOUND on a namespace that
    # carries exactly the attributes those two methods read (the constructor would build the CasADi NLP)
    from ft_mpc.controllers.spiraling_mpc import SpiralingController as RefController
    for cmd in ("hover", "generate_line", "generate_circle", "circle_r_1.5_sPerFullCircle_12"):
        ns = types.SimpleNamespace(Nt=20, dt=dt, mass=model.mass, spiral_params=types.SimpleNamespace(omega_des=np.array([0.0, 0.0, 0.6])))
        RefController.assign_trajectory(ns, np.asarray(load_trajectory(cmd, dt, 3), float))
        out[f"assign::{cmd}::trajectory"] = np.asarray(ns.trajectory, float)
        out[f"assign::{cmd}::nominal_input"] = np.asarray(ns.nominal_input, float)
        xr, ur = RefController.get_next_trajectory_part(ns, 1.7)
        out[f"assign::{cmd}::window_x"], out[f"assign::{cmd}::window_u"] = np.asarray(xr, float), np.asarray(ur, float)
    # ---- export format (controller_debug.py:9-79, 216-260): a synthetic 5-step history through the reference's own
    # DebugVal / ControllerDebug.export; the CSV text is the fixture
    import tempfile
    from ft_mpc.util.controller_debug import ControllerDebug as RefDebug, DebugVal as RefVal
    dbg = RefDebug()
    holder = types.SimpleNamespace(model=model)
    T = 5
    xs = np.concatenate([rng.uniform(-1, 1, (T, 6)), q[:T], rng.uniform(-.5, .5, (T, 3))], axis=1)
    cs = rng.uniform(-1, 1, (T, 13))
    us = rng.uniform(0, model.max_thrust, (T, 16))
    des = rng.uniform(-1, 1, (T, 9))
    for k in range(T):
        dv = RefVal(holder, 0.1 * k)
        dv.set_state(xs[k]); dv.set_circle_state(cs[k]); dv.set_input(us[k], model); dv.set_desired_state(des[k])
        dv.calculate_errors()
        dbg.add_debug_val(dv)
    with tempfile.TemporaryDirectory() as td:
        dbg.export(td + "/dbg")
        out["debug::csv"] = np.frombuffer(Path(td + "/dbg.csv").read_bytes(), dtype=np.uint8)
    out["debug::D"] = np.asarray(model.D, float)
    out["debug::states"], out["debug::centers"], out["debug::thrusts"], out["debug::desired"] = xs, cs, us, des
    dst = ROOT / "tests" / "golden" / "ref_fixtures.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, "with", len(out), "arrays")


if __name__ == "__main__":
    main()
