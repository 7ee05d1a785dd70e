#!/usr/bin/env python
"""Generate tests/golden/ref_nlp_fixtures.npz by EXECUTING THE REFERENCE'S OWN CONTROLLER CODE in the build container.

What runs (unmodified, imported from /root/reference):

    ft_mpc.controllers.spiraling_mpc.SpiralingController.__init__
        -> set_model            (:46-57)   deepcopy(model), InputBounds (Qhull), ControlAllocator.__init__ (CVXPY problem)
        -> set_cost_functions   (:59-85)   running cost ca.Function, load_terminal_ingredients(terminal.yaml)
        -> build_solver         (:87-238)  the NLP  nlp = dict(x, f, g, p), con_lb / con_ub, ca.nlpsol(...)
    SpiralingController.load_trajectory / get_control (:240-317) with a recording `nlpsol` stand-in that returns a
        prescribed decision vector, so that the reference's own post-processing (u_res assembly :301-306) and
        ControlAllocator.get_physical_input (control_allocator.py:65-95) execute on it.

The third-party packages the reference reaches are absent here (casadi, cvxpy, qpsolvers; no network), so the
stand-ins of tools/gen_ref_fixtures.py are extended:
  * casadi        lazy numeric expressions: MX(const), mtimes, reshape (column-major), vcat, vertsplit, fabs, tanh, inf,
                  ca.Function called on expressions (substitution), casadi.tools.struct_symMX / entry, and `nlpsol`,
                  which RECORDS the nlp dict and, when called, returns the decision vector this script prescribes;
  * cvxpy         Parameter / Variable / Minimize / sum_squares / Problem: records objective + constraints as data
                  (the recorded data are the fixture) and solves `min |u|^2 s.t. lin. eq., box` by an exact active-set
                  enumeration on the KKT system (independent of oracle.allocate and of the CUDA allocator).
Nothing of the reference is copied: only numbers leave this script.

Recorded per scenario (faults, horizon, trajectory, time):  z, p = [x0; vec(x_ref); vec(u_ref)], f(z,p), g(z,p), lbg, ubg
at seeded decision vectors (reference ordering z = [u_0..u_{N-1} | x_0..x_N], g = [x_0-x0 | dyn | hull | terminal]);
the allocator's constraint data; and, for the pipeline runs, u_res-derived thrust returned by the reference's get_control.

Usage: python tools/gen_ref_nlp_fixtures.py        (needs /root/reference; run in the build container only)
"""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import gen_ref_fixtures as base  # noqa: E402  (the numeric casadi stand-in)

REF = base.REF
MX = base.MX
_val = base._val


# ---------------------------------------------------------------------------------------------------------
# casadi: what build_solver needs on top of the model subset
# ---------------------------------------------------------------------------------------------------------
def _const(a):
    a = np.asarray(a, dtype=float)
    a = a.reshape(-1, 1) if a.ndim <= 1 else a
    return MX(lambda env: a, a.shape)


class _MXFactory:
    """`ca.MX`: MX.sym / MX.zeros as before, MX(array) / MX(0) -> constant expression"""
    sym = staticmethod(MX.sym)
    zeros = staticmethod(MX.zeros)

    def __call__(self, *a):
        if len(a) == 1:
            return _const(a[0])
        return MX.zeros(*a)

    def __instancecheck__(self, obj):
        return isinstance(obj, MX)


class _SymFunction(base._Function):
    """ca.Function that may also be called on expressions (returns the substituted expression)"""

    def __init__(self, name, ins, outs, opts=None):
        super().__init__(name, ins, outs, opts)
        self.names = []
        for s in ins:                       # recover the symbol names (the sym lambdas look themselves up by name)
            class _Catch(dict):
                def __getitem__(self_inner, k):
                    self.names.append(k)
                    return np.zeros(s.shape)
            s.fn(_Catch())

    def _numeric(self, vals):
        env = {n: np.asarray(v, dtype=complex if np.iscomplexobj(v) else float) for n, v in zip(self.names, vals)}
        return self.outs[0].ev(env)

    def __call__(self, *args):
        if any(isinstance(a, MX) for a in args):
            return MX(lambda env: self._numeric([_val(a, env) for a in args]), self.outs[0].shape)
        return self._numeric([np.asarray(a, dtype=float) for a in args])


class _Struct:
    """casadi.tools.struct_symMX([entry('u', shape=(6,), repeat=N), entry('x', shape=(13,), repeat=N+1)])"""

    def __init__(self, entries):
        flat = []
        for e in entries:                   # the reference passes [(entry, entry)]: one group of entries
            flat.extend(e if isinstance(e[0], tuple) else [e])
        entries = flat
        self.entries = entries
        self.offsets = {}
        o = 0
        for name, shape, rep in entries:
            k = int(np.prod(shape))
            self.offsets[name] = (o, k, rep)
            o += k * rep
        self.size = o
        self.master = MX.sym("opt_var", o, 1)
        self.shape = (o, 1)

    def __getitem__(self, key):
        name, t = key
        o, k, rep = self.offsets[name]
        return self.master[o + t * k: o + (t + 1) * k]

    def __call__(self, value):
        return _NumStruct(self, value)


class _NumStruct:
    def __init__(self, st, value):
        self.st = st
        v = np.asarray(_val(value, {}) if isinstance(value, MX) else value, dtype=float).ravel()
        self.v = np.full(st.size, v[0]) if v.size == 1 else v.copy()

    def __getitem__(self, key):
        if isinstance(key, tuple):
            name, t = key
            o, k, rep = self.st.offsets[name]
            return self.v[o + t * k: o + (t + 1) * k].copy()
        o, k, rep = self.st.offsets[key]
        return [self.v[o + t * k: o + (t + 1) * k].copy() for t in range(rep)]

    def __setitem__(self, key, val):
        if isinstance(key, tuple):
            name, t = key
            o, k, rep = self.st.offsets[name]
            self.v[o + t * k: o + (t + 1) * k] = np.asarray(val, dtype=float).ravel()
        else:
            o, k, rep = self.st.offsets[key]
            for t in range(rep):
                self.v[o + t * k: o + (t + 1) * k] = np.asarray(val[t], dtype=float).ravel()


RECORDED = {}          # the last nlp handed to nlpsol, and what the next solver call shall return


class _Solver:
    def __init__(self, name, plugin, nlp, options):
        self.nlp, self.options, self.plugin = nlp, options, plugin
        RECORDED["nlp"] = nlp
        RECORDED["options"] = dict(options)
        RECORDED["plugin"] = plugin

    def evaluate(self, z, p):
        """f(z, p), g(z, p) of the recorded expressions; p = (x0, x_ref[9,N+1], u_ref[6,N+1])"""
        env = {"opt_var": np.asarray(z).reshape(-1, 1), "x0": np.asarray(p[0]).reshape(-1, 1), "x_ref": np.asarray(p[1]),
               "u_ref": np.asarray(p[2])}
        f = self.nlp["f"].ev(env)
        g = self.nlp["g"].ev(env)
        return float(np.squeeze(f)), np.asarray(g, dtype=float).ravel()

    def __call__(self, x0=None, lbx=None, ubx=None, lbg=None, ubg=None, p=None):
        RECORDED["call"] = dict(x0=np.array(x0.v if isinstance(x0, _NumStruct) else x0, dtype=float).ravel(),
                                p=np.asarray(_val(p, {}), dtype=float).ravel(),
                                lbg=np.asarray(_val(lbg, {}), dtype=float).ravel(), ubg=np.asarray(_val(ubg, {}), dtype=float).ravel())
        z = np.asarray(RECORDED["answer"], dtype=float).ravel()
        pp = RECORDED["call"]["p"]
        N = (z.size - 13) // 19
        pe = (pp[:13], pp[13:13 + 9 * (N + 1)].reshape(N + 1, 9).T, pp[13 + 9 * (N + 1):].reshape(N + 1, 6).T)
        f, _ = self.evaluate(z, pe)
        return {"x": z, "f": f}

    def stats(self):
        return {"return_status": "prescribed_by_fixture_script"}


def _extend_casadi():
    ca = sys.modules["casadi"]
    ca.MX = _MXFactory()
    ca.Function = _SymFunction
    ca.inf = np.inf
    ca.mtimes = lambda a, b: (a @ b) if isinstance(a, MX) else (_const(a) @ b if isinstance(b, MX) else np.asarray(a) @ np.asarray(b))

    def reshape(x, shape):
        def f(env):
            a = x.ev(env)
            return a.reshape(shape, order="F")
        probe = np.zeros(x.shape).reshape(shape, order="F")
        return MX(f, probe.shape)

    def vertcat(*xs):
        if not xs:
            return _const(np.zeros((0, 1)))
        n = sum((x.shape[0] if isinstance(x, MX) else np.size(x)) for x in xs)
        return MX(lambda env: np.vstack([_val(x, env).reshape(-1, 1) for x in xs]), (n, 1))

    ca.reshape = reshape
    ca.vertcat = vertcat
    ca.vcat = lambda xs: vertcat(*xs)
    ca.vertsplit = lambda x: [x[i] for i in range(x.shape[0])]
    ca.fabs = lambda a: MX(lambda env: np.abs(_val(a, env)), a.shape) if isinstance(a, MX) else abs(a)
    ca.tanh = lambda a: MX(lambda env: np.tanh(_val(a, env)), a.shape) if isinstance(a, MX) else np.tanh(a)
    ca.nlpsol = lambda name, plugin, nlp, options: _Solver(name, plugin, nlp, options)
    ct = types.ModuleType("casadi.tools")
    ct.entry = lambda name, shape=(1,), repeat=1: (name, shape, repeat)
    ct.struct_symMX = lambda entries: _Struct(entries)
    sys.modules["casadi.tools"] = ct
    ca.tools = ct
    # MX needs to survive numpy scalar ** and complex evaluation: nothing to add, the lazy ops use numpy throughout


# ---------------------------------------------------------------------------------------------------------
# cvxpy: records the allocation QP, solves it exactly
# ---------------------------------------------------------------------------------------------------------
class _Leaf:
    __array_ufunc__ = None          # `ndarray @ leaf`, `leaf >= ndarray` dispatch to the operators below

    def __init__(self, n, kind):
        self.n, self.kind, self.value = n, kind, None

    def __ge__(self, o): return ("ge", self, o)
    def __le__(self, o): return ("le", self, o)
    def __rmatmul__(self, A): return _Lin(np.asarray(A, dtype=float), self)


class _Lin:
    def __init__(self, A, x):
        self.A, self.x = A, x

    def __eq__(self, o): return ("eq", self, o)


class _Problem:
    def __init__(self, obj, constraints):
        self.obj, self.constraints, self.status = obj, constraints, None
        RECORDED.setdefault("alloc_problems", []).append(self)

    def data(self):
        """(D, d, lb, ub) of  min |u|^2  s.t.  D u = d, lb <= u <= ub  from the recorded expression tuples"""
        kind, var = self.obj
        assert kind == "min_sum_squares"
        n = var.n
        lb, ub, D, d = np.full(n, -np.inf), np.full(n, np.inf), None, None
        for op, lhs, rhs in self.constraints:
            rv = rhs.value if isinstance(rhs, _Leaf) else np.asarray(rhs, dtype=float)
            if op == "ge" and lhs is var:
                lb = np.maximum(lb, rv)
            elif op == "le" and lhs is var:
                ub = np.minimum(ub, rv)
            elif op == "eq" and isinstance(lhs, _Lin) and lhs.x is var:
                D, d = lhs.A, np.asarray(rv, dtype=float)
            else:
                raise NotImplementedError(op)
        return var, D, d, lb, ub

    def solve(self, *a, **k):
        var, D, d, lb, ub = self.data()
        u = box_min_norm(D, d, lb, ub)
        self.status = "optimal" if u is not None else "infeasible"
        var.value = u
        RECORDED.setdefault("alloc_solves", []).append(dict(D=D.copy(), d=d.copy(), lb=lb.copy(), ub=ub.copy(),
                                                            u=None if u is None else u.copy()))
        return None


def box_min_norm(D, d, lb, ub, tol=1e-12):
    """exact solution of  min |u|^2  s.t.  D u = d, lb <= u <= ub  by a primal active-set iteration on the KKT system
    (bounds fixed at lb/ub; multipliers checked).  Small and dense: 16 variables."""
    n = D.shape[1]
    fixed = {}                                   # index -> value
    for _ in range(200):
        free = [i for i in range(n) if i not in fixed]
        uf = np.zeros(n)
        for i, v in fixed.items():
            uf[i] = v
        rhs = d - D @ uf
        Df = D[:, free]
        # min |x|^2 s.t. Df x = rhs  ->  x = Df^+ rhs (minimum norm), multipliers nu from Df Df' nu = rhs
        x, *_ = np.linalg.lstsq(Df, rhs, rcond=None)
        if np.linalg.norm(Df @ x - rhs) > 1e-9 * max(1.0, np.linalg.norm(rhs)):
            x = None
        if x is not None:
            u = uf.copy()
            u[free] = x
            viol = np.maximum(lb - u, 0) + np.maximum(u - ub, 0)
            if viol.max() <= tol:
                # multipliers of the fixed bounds: grad = 2u + D' nu + mu = 0, nu from the free part
                nu, *_ = np.linalg.lstsq(Df.T, -2.0 * x, rcond=None)
                mu = -(2.0 * u + D.T @ nu)
                bad = [i for i, v in fixed.items() if (v == lb[i] and mu[i] > tol) or (v == ub[i] and mu[i] < -tol)]
                bad = [i for i in bad if lb[i] != ub[i]]
                if not bad:
                    return u
                worst = max(bad, key=lambda i: abs(mu[i]))
                del fixed[worst]
                continue
            i = int(np.argmax(viol))
            fixed[i] = lb[i] if u[i] < lb[i] else ub[i]
            continue
        # inconsistent with the current fixing: release the bound with the largest multiplier-like residual
        if not fixed:
            return None
        fixed.pop(next(iter(fixed)))
    return None


def project_onto_polytope(u, G, h):
    """argmin |x - u|^2 s.t. G x <= h  via the dual NNLS  min_{lam >= 0} |G' lam - (u - x_c)|... solved as an active-set
    iteration: x = u - G_A' lam_A with G_A x = h_A, lam_A >= 0 (scipy NNLS on the dual, then an exact re-solve)."""
    from scipy.optimize import nnls
    # dual: min 1/2 |G' lam|^2 - lam'(G u - h), lam >= 0   <=>   NNLS on [G' ; (G u - h)'/s] with the usual lifting
    r = G @ u - h
    if r.max() <= 0.0:
        return u.copy()
    # lifting: |G' lam|^2 - 2 lam' r = |M lam - e|^2 - 1 with M = [G' ; -r'/c ... ] is awkward; iterate instead
    act = [int(np.argmax(r))]
    for _ in range(100):
        GA = G[act]
        lam = np.linalg.solve(GA @ GA.T, GA @ u - h[act])
        if lam.min() < -1e-13:
            act.pop(int(np.argmin(lam)))
            if not act:
                return u.copy()
            continue
        x = u - GA.T @ lam
        rr = G @ x - h
        rr[act] = -np.inf
        if rr.max() <= 1e-13:
            return x
        act.append(int(np.argmax(rr)))
    raise RuntimeError("projection did not converge")


def _install_cvxpy():
    cp = types.ModuleType("cvxpy")
    cp.Parameter = lambda n, **k: _Leaf(n, "param")
    cp.Variable = lambda n, **k: _Leaf(n, "var")
    cp.sum_squares = lambda x: ("sum_squares", x)
    cp.Minimize = lambda e: ("min_" + e[0], e[1])
    cp.Problem = _Problem
    sys.modules["cvxpy"] = cp
    qs = types.ModuleType("qpsolvers")

    def solve_qp(P, q, G, h, solver=None, **k):
        # control_allocator.py:63 hands a 3x3 P with a 6-vector q to a solver that is not in poetry.lock: the branch has
        # no defined result in the reference.  It IS reached at exact KKT points (zero-tolerance membership test :59 with
        # an active hull row at +-1 ulp), so the stand-in returns what the docstring of clip_generalized_input (:43-55)
        # states -- the Euclidean projection of u = -q onto {G x <= h} -- and records how far outside u was.
        u = -np.asarray(q, dtype=float).ravel()
        G, h = np.asarray(G, dtype=float), np.asarray(h, dtype=float).ravel()
        RECORDED.setdefault("clip_calls", []).append(float(np.max(G @ u - h)))
        return project_onto_polytope(u, G, h)
    qs.solve_qp = solve_qp
    sys.modules["qpsolvers"] = qs


# ---------------------------------------------------------------------------------------------------------
def build_reference_controller(faults, N, Q=(1, 1, 1, 1, 1, 1, 2, 2, 2), R=(.1, .1, .1, .01, .01, .01), dt=0.1):
    from ft_mpc.controllers.spiraling_mpc import SpiralingController
    from ft_mpc.models.spiral_model import SpiralModel
    from ft_mpc.models.sys_model import SystemModel
    from ft_mpc.util.broken_thruster import BrokenThruster
    from ft_mpc.util.controller_debug import ControllerDebug
    model = SystemModel(dt)                                           # examples/sim.py:23-29
    for i, a in faults:
        model.set_fault(BrokenThruster(i, a))
    spiral = SpiralModel.from_system_model(model)                     # sim.py:33
    params = {"horizon": N, "param_set": "P1", "P1": {"Q": list(Q), "R": list(R)}}
    ctrl = SpiralingController(spiral, params, ControllerDebug())     # sim.py:41
    return model, spiral, ctrl


SCENARIOS = [
    # tag, faults, N, trajectory command, duration, time of the call
    ("default_N15_hover", [(10, 1.0), (11, 1.0)], 15, "hover", 3, 0.0),
    ("default_N20_circle", [(10, 1.0), (11, 1.0)], 20, "circle_r_1.5_sPerFullCircle_12", 6, 1.3),
    ("dead3_N20_line", [(3, 0.0)], 20, "generate_line", 6, 0.7),
    ("stuck10_N15_circle", [(10, 1.0)], 15, "generate_circle", 6, 2.1),
    ("dead0dead5_N15_hover", [(0, 0.0), (5, 0.0)], 15, "hover_1_-2_0.5", 3, 0.4),
]


def main():
    if not REF.exists():
        raise SystemExit("/root/reference is not available: fixtures can only be generated in the build container")
    base._install_stubs()
    _extend_casadi()
    _install_cvxpy()
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT / "oracle"))
    import ftmpc_oracle as o                                          # only to prescribe KKT points for the pipeline runs
    from scipy.spatial.transform import Rotation

    out = {}
    rng = np.random.default_rng(77)
    tags = []
    for tag, faults, N, cmd, dur, tcall in SCENARIOS:
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            model, spiral, ctrl = build_reference_controller(faults, N)
            ctrl.load_trajectory(cmd, dur)                            # spiraling_mpc.py:240-253
        solver = ctrl.solver
        tags.append(tag)
        out[f"{tag}::faults"] = np.array(faults, dtype=float).reshape(-1, 2)
        out[f"{tag}::N"] = np.array(N)
        out[f"{tag}::options"] = np.array(sorted(f"{k}={v}" for k, v in RECORDED["options"].items()))
        out[f"{tag}::plugin"] = np.array(RECORDED["plugin"])
        out[f"{tag}::num_var"] = np.array(ctrl.num_var)
        out[f"{tag}::u_comp"] = np.asarray(ctrl.u_comp, float).ravel()
        out[f"{tag}::lbg"] = np.asarray(_val(ctrl.con_lb, {}), float).ravel()
        out[f"{tag}::ubg"] = np.asarray(_val(ctrl.con_ub, {}), float).ravel()
        out[f"{tag}::lbx"] = ctrl.optvar_lb.v
        out[f"{tag}::ubx"] = ctrl.optvar_ub.v
        xr, ur = ctrl.get_next_trajectory_part(tcall)                  # :356-365
        out[f"{tag}::x_ref"], out[f"{tag}::u_ref"] = np.asarray(xr, float), np.asarray(ur, float)
        # ---- f, g at seeded decision vectors
        K = 3
        zs, x0s, fs, gs = [], [], [], []
        for k in range(K):
            q = Rotation.random(N + 1, random_state=100 * len(tags) + k).as_quat()
            X = np.concatenate([rng.uniform(-1, 1, (N + 1, 3)), rng.uniform(-.5, .5, (N + 1, 3)),
                                rng.uniform(-.5, .5, (N + 1, 3)) + [0, 0, .6], q * rng.uniform(0.9, 1.1, (N + 1, 1))], axis=1)
            U = rng.normal(0, 2.0, (N, 6))
            z = np.concatenate([U.ravel(), X.ravel()])
            x0 = X[0] + rng.normal(0, 0.05, 13)
            f, g = solver.evaluate(z, (x0, xr, ur))
            zs.append(z); x0s.append(x0); fs.append(f); gs.append(g)
        out[f"{tag}::z"], out[f"{tag}::x0"] = np.array(zs), np.array(x0s)
        out[f"{tag}::f"], out[f"{tag}::g"] = np.array(fs), np.array(gs)
        # ---- the allocator's problem as the reference's ControlAllocator.__init__ stated it (control_allocator.py:20-40)
        prob = ctrl.contr_alloc.prob
        ctrl.contr_alloc.u_desired.value = np.zeros(6)
        ctrl.contr_alloc.upper_bound.value = np.asarray(ctrl.model.u_ub_physical, float).ravel()
        var, D, d, lb, ub = prob.data()
        out[f"{tag}::alloc_D"], out[f"{tag}::alloc_lb"], out[f"{tag}::alloc_ub"] = D, lb, ub
        out[f"{tag}::alloc_objective"] = np.array(prob.obj[0])
        # ---- pipeline: the reference's get_control with the solver returning the oracle's KKT point
        robot = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-.5, .5, 3), Rotation.random(random_state=5 + len(tags)).as_quat(),
                                np.array([0, 0, .6]) + rng.uniform(-.3, .3, 3)])
        c0 = np.asarray(spiral.robot_to_center(robot), float).ravel()
        fsr = o.FaultSet([(int(i), float(a)) for i, a in faults])
        prob_o = o.Problem(fsr, N, c0, np.asarray(xr, float).T.copy(), np.asarray(ur, float).T.copy())
        sol = o.solve_nlp(prob_o)
        zstar = np.concatenate([np.asarray(sol["U"]).ravel(), np.asarray(sol["X"]).ravel()])
        RECORDED["answer"] = zstar
        RECORDED["alloc_solves"] = []
        RECORDED["clip_calls"] = []
        ctrl.optimal_solution = None
        with contextlib.redirect_stdout(io.StringIO()):
            thrust = ctrl.get_control(robot, tcall)                   # :288-317
        call = RECORDED["call"]
        rec = RECORDED["alloc_solves"][-1]
        out[f"{tag}::pipe_robot"], out[f"{tag}::pipe_c0"], out[f"{tag}::pipe_z"] = robot, c0, zstar
        out[f"{tag}::pipe_p"], out[f"{tag}::pipe_x0guess"] = call["p"], call["x0"]
        out[f"{tag}::pipe_udes"], out[f"{tag}::pipe_ub"] = rec["d"], rec["ub"]
        out[f"{tag}::pipe_thrust"] = np.asarray(thrust, float).ravel()
        # > 0 when the reference's zero-tolerance membership test sent the KKT point into the (undefined) clip branch
        out[f"{tag}::pipe_clip_excess"] = np.array(RECORDED["clip_calls"][:1] or [0.0])
        out[f"{tag}::pipe_kkt"] = np.array([sol["kkt_stat"], sol["kkt_viol"]])
        # second call: the warm start the reference builds from its previous solution (:324-334)
        with contextlib.redirect_stdout(io.StringIO()):
            ctrl.get_control(robot, tcall)
        out[f"{tag}::pipe_x0guess_warm"] = RECORDED["call"]["x0"]
        print(tag, "nz", ctrl.num_var, "ng", out[f"{tag}::lbg"].size, "f", fs[0], "thrust", np.round(thrust, 4))
    out["tags"] = np.array(tags)
    dst = ROOT / "tests" / "golden" / "ref_nlp_fixtures.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, "with", len(out), "arrays")


if __name__ == "__main__":
    main()
