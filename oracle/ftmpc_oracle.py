"""CPU oracle for the ft_mpc per-timestep MPC solve  --  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product package (`fault-tolerant-mpc_b200/`, imported as `ft_mpc_b200`)
never does; it fails loudly when its CUDA library is missing.

PARITY: the problem definition, the NLP (f, g, lbg, ubg, row order), the allocation QP and the get_control
post-processing are PINNED against the reference's own code; which local solution IPOPT (tol 1e-3) would return is not.
  * tools/gen_ref_fixtures.py executes ft_mpc.util.utils, ft_mpc.models.sys_model / spiral_model, SpiralParameters,
    InputBounds and get_trajectory in the build container (numeric stand-in for the few CasADi calls they make) and
    tests/test_reference_fixtures.py checks every model function below against those outputs
    (tests/golden/ref_fixtures.npz) -- including the ROW ORDER of the input-bound hull, i.e. the numbering of the
    constraints -- and, likewise executed from the reference, SpiralingController.assign_trajectory /
    get_next_trajectory_part (reference window + nominal wrench) and the CSV written by ControllerDebug.export.
    The stored terminal ingredients (config/terminal.yaml) are pinned through sympy evaluation
    (tools/gen_terminal_data.py).
  * tools/gen_ref_nlp_fixtures.py executes the reference's SpiralingController.__init__ -> set_model /
    set_cost_functions / build_solver (spiraling_mpc.py:27-238), ControlAllocator.__init__ / get_physical_input
    (control_allocator.py:12-95) and get_control (:288-317) on numeric casadi / cvxpy stand-ins and records f(z,p),
    g(z,p), lbg, ubg of the reference's `nlp` dict at seeded decision vectors (non-zero u_ref included), the
    allocation QP's data and the thrust the reference's own post-processing returns for a prescribed NLP solution;
    tests/test_reference_nlp.py holds Problem.nlp_eval, allocate and get_control below to them (1e-12).
  * Not pinnable here: the reference ships no tests or golden vectors for solve_mpc, and casadi 3.6.7 / IPOPT and
    cvxpy 1.6.4 / OSQP are not installable offline, so the optimiser OUTPUT (KKT point, active set) comes from an
    independent solver on the (pinned) NLP: scipy SLSQP + Newton polish to tight KKT tolerance.  The NLP is
    non-convex; "same local minimum as IPOPT" rests on the survey's probe and stays a stated assumption.
Every function cites the reference file:line it follows (paths relative to /root/reference).

Derivatives in this file are obtained by complex-step differentiation (the prediction model is a
polynomial map, so complex-step is exact to rounding) -- deliberately independent of the
hand-derived Jacobians inside the CUDA kernels.
"""
from __future__ import annotations

import itertools
import json
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

# --------------------------------------------------------------------------------------------
# constants                                              ft_mpc/models/sys_model.py:52-131
# --------------------------------------------------------------------------------------------
MASS = 16.8
INERTIA = np.diag([0.2, 0.3, 0.25])
INERTIA_INV = np.linalg.inv(INERTIA)
MAX_THRUST = 3.4
NX, NU, NTHR, NOPT = 13, 6, 16, 9


def allocation_matrix() -> np.ndarray:
    """D (6x16): rows [Fx,Fy,Fz,tx,ty,tz], columns = thrusters.  sys_model.py:73-123"""
    D = np.zeros((NU, NTHR))
    d1, d2, d3 = 0.12, 0.09, 0.05
    D[0, [0, 1, 4, 5]] = -1.0
    D[0, [2, 3, 6, 7]] = 1.0
    D[1, [8, 9]] = -1.0
    D[1, [10, 11]] = 1.0
    D[2, [12, 14]] = -1.0
    D[2, [13, 15]] = 1.0
    D[3, [12, 15]] = -d1
    D[3, [13, 14]] = d1
    D[4, [0, 3, 4, 7]] = -d3
    D[4, [1, 2, 5, 6]] = d3
    D[5, [0, 1, 6, 7]] = d1
    D[5, [2, 3, 4, 5]] = -d1
    D[5, [8, 11]] = -d2
    D[5, [9, 10]] = d2
    return D


D_ALLOC = allocation_matrix()

# spiral parameters                         controllers/tools/spiral_parameters.py:21,26-39
OMEGA_DES = np.array([0.0, 0.0, 0.6])
F_VIRT = 3.5 * np.array([0.0, 1.0, 0.0])
R_VEC = np.linalg.norm(F_VIRT) / (MASS * np.linalg.norm(OMEGA_DES) ** 2) * np.array([0.0, 1.0, 0.0])
BETA = np.array([0.0, 0.0, 0.0, 1.0])       # identity quaternion => every RotFull(beta) is I

# default tuning                                         ft_mpc/config/reactive.yaml:25-41
Q_DEFAULT = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2], dtype=float)
R_DEFAULT = np.array([0.1, 0.1, 0.1, 0.01, 0.01, 0.01], dtype=float)

_DATA = Path(__file__).resolve().parent.parent / "fault-tolerant-mpc_b200" / "data"


# --------------------------------------------------------------------------------------------
# faults                                  sys_model.py:228-243, util/broken_thruster.py:1-10
# --------------------------------------------------------------------------------------------
@dataclass
class FaultSet:
    """A list of (thruster index, intensity in [0,1]) pairs, as appended by SystemModel.set_fault."""
    faults: list = field(default_factory=list)

    @property
    def faulty_force(self) -> np.ndarray:
        f = np.zeros(NTHR)
        for i, inten in self.faults:
            f[i] = inten * MAX_THRUST                      # sys_model.py:239
        return f

    @property
    def ub(self) -> np.ndarray:
        ub = np.full(NTHR, MAX_THRUST)
        for i, _ in self.faults:
            ub[i] = 0.0                                    # sys_model.py:240
        return ub

    @property
    def generalized(self) -> np.ndarray:
        return D_ALLOC @ self.faulty_force                 # sys_model.py:241

    @property
    def mask(self) -> int:
        m = 0
        for i, _ in self.faults:
            m |= 1 << i
        return m

    @property
    def u_comp(self) -> np.ndarray:
        """compensation_force, spiral_parameters.py:37 (RobotToCenterRot(beta) = I, spiraling_mpc.py:141)"""
        return np.concatenate([F_VIRT, np.zeros(3)]) - self.generalized


# --------------------------------------------------------------------------------------------
# rotation helpers                                                    util/utils.py:4-74
# --------------------------------------------------------------------------------------------
def rot(q):
    """Rot(q), q=[x,y,z,w] (world->body, NOT normalised).  utils.py:4-19.  Batched over leading dims."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    r0 = np.stack([x * x - y * y - z * z + w * w, 2 * (x * y + z * w), 2 * (x * z - y * w)], -1)
    r1 = np.stack([2 * (x * y - z * w), -x * x + y * y - z * z + w * w, 2 * (y * z + x * w)], -1)
    r2 = np.stack([2 * (x * z + y * w), 2 * (y * z - x * w), -x * x - y * y + z * z + w * w], -1)
    return np.stack([r0, r1, r2], -2)


def rot_inv_apply(q, v):
    """RotInv(q) @ v = Rot(q)^T v   (utils.py:21-31)"""
    return np.einsum("...ji,...j->...i", rot(q), v)


def omega_apply(w, q):
    """OmegaOperator(w) @ q.  sys_model.py:8-29"""
    wx, wy, wz = w[..., 0], w[..., 1], w[..., 2]
    x, y, z, s = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([
        wz * y - wy * z + wx * s,
        -wz * x + wx * z + wy * s,
        wy * x - wx * y + wz * s,
        -wx * x - wy * y - wz * z,
    ], -1)


def cross(a, b):
    return np.stack([
        a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0],
    ], -1)


# --------------------------------------------------------------------------------------------
# prediction model (orbit-centre state)  models/spiral_model.py:44-76, sys_model.py:138-162
# --------------------------------------------------------------------------------------------
def spiral_dxdt(c, u, df):
    """SpiralModel.dx_dt.  c=[p_c,v_c,omega,q] (13), u generalised wrench (6), df = D f_fault (6)."""
    vel, omega, q = c[..., 3:6], c[..., 6:9], c[..., 9:13]
    gen = u + df                                                         # spiral_model.py:61
    force, torque = gen[..., 0:3], gen[..., 3:6]
    jdiag = np.diag(INERTIA)
    io = omega * jdiag                                                   # :65
    domega = (torque - cross(omega, io)) / jdiag                         # :66-67
    rr = np.broadcast_to(R_VEC, omega.shape)
    body = force / MASS + cross(domega, rr) + cross(omega, cross(omega, rr))
    dvel = rot_inv_apply(q, body)                                        # :69-73  RotCasadi(q).T @ (...)
    dq = 0.5 * omega_apply(omega, q)                                     # :75
    return np.concatenate([vel, dvel, domega, dq], -1)                   # :76


def spiral_rk4(c, u, df, dt):
    """SystemModel.rk4_integrator applied to SpiralModel.dx_dt.  sys_model.py:150-158 (no quaternion renormalisation)"""
    k1 = spiral_dxdt(c, u, df)
    k2 = spiral_dxdt(c + dt / 2 * k1, u, df)
    k3 = spiral_dxdt(c + dt / 2 * k2, u, df)
    k4 = spiral_dxdt(c + dt * k3, u, df)
    return c + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


def robot_to_center(x):
    """SpiralModel.robot_to_center: robot [p,v,q,omega] -> centre [p_c,v_c,omega,q].  spiral_model.py:91-109"""
    omega, q = x[..., 10:13], x[..., 6:10]
    rr = np.broadcast_to(R_VEC, omega.shape)
    pos = x[..., 0:3] + rot_inv_apply(q, rr)
    vel = x[..., 3:6] + rot_inv_apply(q, cross(omega, rr))
    return np.concatenate([pos, vel, omega, q], -1)


# --------------------------------------------------------------------------------------------
# plant (16 thrusters, robot state order [p,v,q,omega])           sys_model.py:177-226
# --------------------------------------------------------------------------------------------
def plant_dxdt(x, u, fs: FaultSet):
    vel, q, omega = x[..., 3:6], x[..., 6:10], x[..., 10:13]
    u = np.where(fs.ub > 0.0, u, 0.0)                                    # :198-206 failed inputs zeroed
    gen = (u + fs.faulty_force) @ D_ALLOC.T                              # :211
    force, torque = gen[..., 0:3], gen[..., 3:6]
    dv = rot_inv_apply(q, force) / MASS                                  # :219
    dq = 0.5 * omega_apply(omega, q)                                     # :222
    jdiag = np.diag(INERTIA)
    domega = (torque - cross(omega, omega * jdiag)) / jdiag              # :225-227
    return np.concatenate([vel, dv, dq, domega], -1)


def plant_rk4(x, u, fs: FaultSet, dt):
    k1 = plant_dxdt(x, u, fs)
    k2 = plant_dxdt(x + dt / 2 * k1, u, fs)
    k3 = plant_dxdt(x + dt / 2 * k2, u, fs)
    k4 = plant_dxdt(x + dt * k3, u, fs)
    return x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


def normalize_quaternion_robot(x):
    """SystemModel.normalize_quaternion.  sys_model.py:164-175"""
    x = np.array(x, dtype=float)
    x[..., 6:10] = x[..., 6:10] / np.linalg.norm(x[..., 6:10], axis=-1, keepdims=True)
    return x


# --------------------------------------------------------------------------------------------
# input-bound polytope                         controllers/tools/input_bounds.py:43-76
# --------------------------------------------------------------------------------------------
_HULL_CACHE: dict = {}


def input_bounds(fs: FaultSet):
    """InputBounds.calc_input_bounds: Qhull of the 2^(#healthy) corner wrenches, np.unique on the
    facet equations.  Returns (A_h [n_h,6], b_h [n_h])  (rows sorted lexicographically by np.unique)."""
    key = tuple(sorted((int(i), float(a)) for i, a in fs.faults))
    if key in _HULL_CACHE:
        return _HULL_CACHE[key]
    from scipy.spatial import ConvexHull
    broken = [i for i, _ in fs.faults]
    ff = fs.faulty_force
    min_max = []
    for i in range(NTHR):
        min_max.append([ff[i], ff[i]] if i in broken else [0.0, MAX_THRUST])     # :49-55
    # itertools.product(*min_max), then ONE matrix-vector product per corner exactly as the reference does (:57-63).
    # The per-corner np.matmul matters: a vectorised `corners @ D.T` rounds differently in the last bit, Qhull's facet
    # normals inherit that noise, and np.unique (:71) then sorts the rows -- i.e. numbers the constraints --
    # differently from the reference (checked against tests/golden/ref_fixtures.npz).
    corners = np.array(list(itertools.product(*min_max)))
    verts = np.array([np.matmul(D_ALLOC, c) for c in corners])
    verts = np.unique(verts, axis=0)                                              # :67
    hull = ConvexHull(verts)                                                      # :68
    simplified = np.unique(hull.equations, axis=0)                                # :71
    A, b = simplified[:, :-1], -simplified[:, -1]                                 # :72-73
    _HULL_CACHE[key] = (A, b)
    return A, b


# --------------------------------------------------------------------------------------------
# terminal ingredients        controllers/tools/terminal_ingredients.py:451-474 + config/terminal.yaml
# --------------------------------------------------------------------------------------------
class TerminalIngredients:
    """Numeric table derived from terminal.yaml by tools/gen_terminal_data.py (sympy parse, no eval)."""

    def __init__(self, path=None):
        d = json.loads(Path(path or _DATA / "terminal.json").read_text())
        self.const = d["const"]
        self.poly = [(t["coeff"], np.array(t["exps"])) for t in d["poly"]]
        self.root = [(t["coeff"], np.array(t["exps"]), t["eps"], t["pow"]) for t in d["root"]]
        self.A = np.array(d["A"])
        self.b = np.array(d["b"])
        self.anchors = d["anchors"]

    def cost(self, e):
        """V_f(e), e = x_N[0:9] - xr_N.  Works for complex e (complex-step differentiation)."""
        e = np.asarray(e)
        v = self.const + 0 * e[..., 0]
        for c, p in self.poly:
            v = v + c * np.prod(e ** p, axis=-1)
        for c, p, eps, w in self.root:
            v = v + c * (np.prod(e ** p, axis=-1) + eps) ** w
        return v

    def grad(self, e):
        h = 1e-30
        g = np.zeros(9)
        for i in range(9):
            ec = np.array(e, dtype=complex)
            ec[i] += 1j * h
            g[i] = self.cost(ec).imag / h
        return g


TERMINAL = TerminalIngredients()


# --------------------------------------------------------------------------------------------
# the NLP                                   controllers/spiraling_mpc.py:87-238 (build_solver)
# --------------------------------------------------------------------------------------------
@dataclass
class Problem:
    """One MPC instance: everything `build_solver` + `get_control` see for one call."""
    fs: FaultSet
    N: int
    c0: np.ndarray                     # centre state (13)
    xref: np.ndarray                   # (N+1, 9)   columns of x_ref, spiraling_mpc.py:104,356-365
    uref: np.ndarray                   # (N+1, 6)
    Q: np.ndarray = field(default_factory=lambda: Q_DEFAULT.copy())
    R: np.ndarray = field(default_factory=lambda: R_DEFAULT.copy())
    dt: float = 0.1

    def __post_init__(self):
        self.A_h, self.b_h = input_bounds(self.fs)
        self.n_h = self.A_h.shape[0]
        self.df = self.fs.generalized
        self.u_comp = self.fs.u_comp
        self.u_unc = self.df.copy()              # spiraling_mpc.py:145

    # ---- pieces of the stage map -------------------------------------------------------
    def u_ref_rot(self, x_t, t):
        """[Rot(q_t)^T ur[0:3]; ur[3:6]]   spiraling_mpc.py:156-166"""
        ur = self.uref[t]
        return np.concatenate([rot_inv_apply(x_t[..., 9:13], np.broadcast_to(ur[0:3], x_t[..., 0:3].shape)),
                               np.broadcast_to(ur[3:6], x_t[..., 0:3].shape)], -1)

    def stage(self, x_t, u_t, t):
        """x_{t+1} and hull lhs for stage t.  spiraling_mpc.py:171,175"""
        urr = self.u_ref_rot(x_t, t)
        x_next = spiral_rk4(x_t, u_t + urr + self.u_comp, self.df, self.dt)
        hull = (u_t + urr + self.u_comp + self.u_unc) @ self.A_h.T
        return x_next, hull

    def rollout(self, U):
        U = np.asarray(U).reshape(self.N, NU)
        X = np.zeros((self.N + 1, NX), dtype=U.dtype)
        X[0] = self.c0
        hull = np.zeros((self.N, self.n_h), dtype=U.dtype)
        for t in range(self.N):
            X[t + 1], hull[t] = self.stage(X[t], U[t], t)
        return X, hull

    def objective_from(self, X, U):
        """obj, spiraling_mpc.py:188,195-196"""
        e = X[: self.N, 0:9] - self.xref[: self.N]
        obj = np.sum(e * e * self.Q) + np.sum(U * U * self.R)
        return obj + TERMINAL.cost(X[self.N, 0:9] - self.xref[self.N])

    # ---- literal multiple-shooting restatement (for checking a candidate z) -------------
    def nlp_eval(self, z):
        """f(z) and g(z) in the reference's ordering.  z = [u_0..u_{N-1} | x_0..x_N]  (:110-114)
        g = [x_0 - x0 | dyn_0.. | hull_0.. | terminal]  (:206-214).  Returns f, g, lbg, ubg."""
        N = self.N
        U = z[: NU * N].reshape(N, NU)
        X = z[NU * N:].reshape(N + 1, NX)
        eq = [X[0] - self.c0]
        hull = []
        for t in range(N):
            xn, h = self.stage(X[t], U[t], t)
            eq.append(xn - X[t + 1])
            hull.append(h)
        term = TERMINAL.A @ (X[N, 0:9] - self.xref[N])
        g = np.concatenate(eq + hull + [term])
        neq = NX * (N + 1)
        lbg = np.concatenate([np.zeros(neq), np.full(self.n_h * N + 72, -np.inf)])
        ubg = np.concatenate([np.zeros(neq), np.tile(self.b_h, N), TERMINAL.b])
        return self.objective_from(X, U), g, lbg, ubg

    # ---- reduced (single-shooting) problem: same NLP with the equality rows eliminated ---
    def ineq(self, U):
        """c(U) <= 0 in the reference's inequality order: hull_0..hull_{N-1}, terminal."""
        X, hull = self.rollout(U)
        term = TERMINAL.A @ (X[self.N, 0:9] - self.xref[self.N]) - TERMINAL.b
        return np.concatenate([(hull - self.b_h).ravel(), term]), X

    def linearize(self, X, U):
        """A_t = d x_{t+1}/d x_t, B_t = d x_{t+1}/d u_t, and d hull_t / d x_t by complex step."""
        N, h = self.N, 1e-30
        U = np.asarray(U).reshape(N, NU)
        A = np.zeros((N, NX, NX)); B = np.zeros((N, NX, NU)); Hx = np.zeros((N, self.n_h, NX))
        for t in range(N):
            xp = np.tile(X[t].astype(complex), (NX + NU, 1))
            up = np.tile(U[t].astype(complex), (NX + NU, 1))
            xp[np.arange(NX), np.arange(NX)] += 1j * h
            up[NX + np.arange(NU), np.arange(NU)] += 1j * h
            xn, hl = self.stage(xp, up, t)
            A[t] = xn[:NX].imag.T / h
            B[t] = xn[NX:].imag.T / h
            Hx[t] = hl[:NX].imag.T / h
        return A, B, Hx

    def sensitivities(self, X, U):
        """G[t] = d x_t / d U  ((N+1) x 13 x 6N) by the forward recursion."""
        N = self.N
        A, B, Hx = self.linearize(X, U)
        G = np.zeros((N + 1, NX, NU * N))
        for t in range(N):
            G[t + 1] = A[t] @ G[t]
            G[t + 1][:, NU * t: NU * (t + 1)] += B[t]
        return G, Hx

    def fun_and_grad(self, U):
        """objective, gradient, constraints c(U) (<=0 feasible) and Jacobian dc/dU."""
        N = self.N
        U = np.asarray(U, dtype=float).reshape(N, NU)
        X, hull = self.rollout(U)
        G, Hx = self.sensitivities(X, U)
        e = X[:N, 0:9] - self.xref[:N]
        eN = X[N, 0:9] - self.xref[N]
        f = np.sum(e * e * self.Q) + np.sum(U * U * self.R) + float(TERMINAL.cost(eN))
        grad = (2 * U * self.R).ravel()
        for t in range(N):
            grad += G[t][0:9].T @ (2 * self.Q * e[t])
        grad += G[N][0:9].T @ TERMINAL.grad(eN)
        c = np.concatenate([(hull - self.b_h).ravel(), TERMINAL.A @ eN - TERMINAL.b])
        J = np.zeros((self.n_h * N + 72, NU * N))
        for t in range(N):
            rows = slice(self.n_h * t, self.n_h * (t + 1))
            J[rows] = Hx[t] @ G[t]
            J[rows, NU * t: NU * (t + 1)] += self.A_h
        J[self.n_h * N:] = TERMINAL.A @ G[N][0:9]
        return f, grad, c, J, X


ACTIVE_TOL = 1e-7       # row i active  <=>  b_i - g_i(z*) <= ACTIVE_TOL   


def kkt_residual(prob: Problem, U):
    """Solver-independent optimality check of a candidate U for the reference NLP.
    Multipliers: non-negative least squares on the (near-)active rows.  Returns a dict with
    stationarity residual (inf-norm, relative to max(1,|grad|)), max violation, active rows."""
    from scipy.optimize import nnls
    f, grad, c, J, X = prob.fun_and_grad(U)
    act = np.where(c >= -1e-6)[0]
    if len(act):
        lam, _ = nnls(J[act].T, -grad, maxiter=20 * len(act) + 200)
        r = grad + J[act].T @ lam
    else:
        lam, r = np.zeros(0), grad
    return {"f": f, "stat": float(np.max(np.abs(r)) / max(1.0, np.max(np.abs(grad)))),
            "viol": float(max(0.0, c.max())), "active": np.where(c >= -ACTIVE_TOL)[0],
            "lam_rows": act, "lam": lam, "c": c, "X": X}


def polish_newton(prob: Problem, U, iters=6, h=1e-6, tol=1e-11):
    """Newton iteration on the KKT equations [grad f + J_A^T lam ; c_A] = 0 for the active rows A
    identified at U (Lagrangian Hessian by central differences of the complex-step gradient).
    Independent of the GPU's SQP: no QP sub-solver, no Gauss-Newton model.  Returns (U, lam, A, ok)."""
    U = np.asarray(U, dtype=float).ravel().copy()
    n = U.size
    k = kkt_residual(prob, U)
    act = k["active"]
    lam_full = dict(zip(k["lam_rows"], k["lam"]))
    lam = np.array([lam_full.get(i, 0.0) for i in act])

    def grad_l(Uv, lv):
        f, g, c, J, X = prob.fun_and_grad(Uv)
        return (g + J[act].T @ lv if len(act) else g), c[act], J[act], c

    ok = False
    for _ in range(iters):
        gl, ca, JA, c = grad_l(U, lam)
        res = max(np.abs(gl).max(), np.abs(ca).max() if len(act) else 0.0)
        if res < tol:
            ok = True
            break
        H = np.zeros((n, n))
        for i in range(n):
            e = np.zeros(n); e[i] = h
            H[:, i] = (grad_l(U + e, lam)[0] - grad_l(U - e, lam)[0]) / (2 * h)
        H = 0.5 * (H + H.T)
        m = len(act)
        K = np.block([[H, JA.T], [JA, np.zeros((m, m))]])
        try:
            d = np.linalg.solve(K, -np.concatenate([gl, ca]))
        except np.linalg.LinAlgError:
            break
        if not np.all(np.isfinite(d)) or np.abs(d[:n]).max() > 0.5:
            break
        U = U + d[:n]
        lam = lam + d[n:]
    if ok:
        c = prob.ineq(U)[0]
        inactive = np.setdiff1d(np.arange(c.size), act)
        ok = bool((lam.min() if len(lam) else 1.0) > -1e-9 and (c[inactive].max() if len(inactive) else -1.0) < 1e-9)
    return U, lam, act, ok


def solve_nlp(prob: Problem, U0=None, ftol=1e-14, maxiter=600, polish=True):
    """Solve the reference NLP (reduced form) with scipy SLSQP from U0 (default: zeros, i.e. the
    u=0 forward rollout standing in for the reference's all-zero cold start, spiraling_mpc.py:331-334),
    then (polish=True) tighten with `polish_newton`, then report the KKT residual.
    Independent of the SQP/QP algorithm used on the GPU."""
    from scipy.optimize import minimize
    N = prob.N
    cache = {}

    def ev(U):
        k = U.tobytes()
        if k not in cache:
            cache.clear()
            cache[k] = prob.fun_and_grad(U)
        return cache[k]

    U0 = np.zeros(NU * N) if U0 is None else np.asarray(U0, dtype=float).ravel()
    with np.errstate(all="ignore"):
        res = minimize(lambda U: ev(U)[0], U0, jac=lambda U: ev(U)[1], method="SLSQP",
                       constraints=[{"type": "ineq", "fun": lambda U: -ev(U)[2], "jac": lambda U: -ev(U)[3]}],
                       options={"ftol": ftol, "maxiter": maxiter})
    U = res.x
    polished = False
    if polish and np.all(np.isfinite(U)):
        k0 = kkt_residual(prob, U)
        if k0["viol"] < 1e-6 and k0["stat"] < 1e-3:
            U2, lam, act, polished = polish_newton(prob, U)
            if polished:
                U = U2
    k = kkt_residual(prob, U)
    X = k["X"]
    return {"U": U.reshape(N, NU), "X": X, "f": k["f"], "active": k["active"], "kkt_stat": k["stat"],
            "kkt_viol": k["viol"], "nit": res.nit, "status": int(res.status), "message": res.message,
            "c": k["c"], "polished": polished}


# --------------------------------------------------------------------------------------------
# output assembly + control allocation     spiraling_mpc.py:301-307, control_allocator.py:28-95
# --------------------------------------------------------------------------------------------
def assemble_u_res(prob: Problem, u0):
    """u_res = u*_0 + RotFullInv(q_0) ur_0 + u_comp  (then RotFull(beta)=I).  spiraling_mpc.py:301-306"""
    ur = prob.uref[0]
    return u0 + np.concatenate([rot_inv_apply(prob.c0[9:13], ur[0:3]), ur[3:6]]) + prob.u_comp


def clip_generalized_input(prob: Problem, u, tol=1e-9):
    """ControlAllocator.clip_generalized_input (control_allocator.py:42-63).  The reference's
    projection branch cannot run (3x3 P vs 6-D u, solver 'daqp' not installed); the restated semantics
    are: return u if A_h u <= b_h + tol, else the Euclidean projection onto {A_h u <= b_h}."""
    if np.all(prob.A_h @ u <= prob.b_h + tol):
        return u
    from scipy.optimize import minimize
    res = minimize(lambda v: 0.5 * np.sum((v - u) ** 2), u, jac=lambda v: v - u, method="SLSQP",
                   constraints=[{"type": "ineq", "fun": lambda v: prob.b_h - prob.A_h @ v,
                                 "jac": lambda v: -prob.A_h}], options={"ftol": 1e-16, "maxiter": 200})
    return res.x


def allocate(u_des, ub):
    """min ||u||^2  s.t.  D u = u_des, 0 <= u <= ub   (control_allocator.py:28-40).
    The reference hands this to CVXPY/OSQP (eps 1e-5); here: scipy SLSQP, then an exact polish --
    bounds within 1e-7 are taken as active and the least-norm KKT system on the free set is solved.
    Returns (u_phys[16], ok)."""
    from scipy.optimize import minimize
    D = D_ALLOC
    n = NTHR
    free0 = [i for i in range(n) if ub[i] > 0.0]
    bnds = [(0.0, float(ub[i])) for i in range(n)]
    res = minimize(lambda u: float(u @ u), np.clip(np.linalg.pinv(D) @ u_des, 0.0, ub), jac=lambda u: 2 * u,
                   method="SLSQP", bounds=bnds,
                   constraints=[{"type": "eq", "fun": lambda u: D @ u - u_des, "jac": lambda u: D}],
                   options={"ftol": 1e-16, "maxiter": 500})
    u = np.clip(res.x, 0.0, ub)
    lo = [i for i in free0 if u[i] <= 1e-7]
    hi = [i for i in free0 if u[i] >= ub[i] - 1e-7]
    free = [i for i in free0 if i not in lo and i not in hi]
    up = np.zeros(n)
    up[hi] = ub[hi]
    rhs = u_des - D @ up
    if free:
        Df = D[:, free]
        y, *_ = np.linalg.lstsq(Df @ Df.T, rhs, rcond=None)
        up[free] = Df.T @ y
    ok = (np.linalg.norm(D @ up - u_des) <= 1e-8 * max(1.0, np.linalg.norm(u_des))
          and up.min() >= -1e-9 and np.all(up <= ub + 1e-9))
    if ok and np.abs(up - u).max() < 1e-4:
        return np.clip(up, 0.0, ub), True
    ok = np.linalg.norm(D @ u - u_des) <= 1e-6 * max(1.0, np.linalg.norm(u_des))
    return u, bool(ok)


def get_control(prob: Problem, sol=None):
    """Everything SpiralingController.get_control returns for one call (spiraling_mpc.py:288-317)."""
    sol = sol or solve_nlp(prob)
    u_res = assemble_u_res(prob, sol["U"][0])
    u_fault = prob.df
    u_des = clip_generalized_input(prob, u_res + u_fault) - u_fault           # control_allocator.py:79
    thrust, ok = allocate(u_des, prob.fs.ub)
    return {"u0": sol["U"][0].copy(), "u_res": u_res, "u_des": u_des, "thrust": thrust, "alloc_ok": ok, **sol}


# --------------------------------------------------------------------------------------------
# reference window                           spiraling_mpc.py:255-286, 356-365
# --------------------------------------------------------------------------------------------
def assign_trajectory(traj, N, dt):
    """traj: 13 x T robot-state reference.  Returns (trajectory 9 x (T+N), nominal_input 6 x (T+N))."""
    orig = np.hstack((traj, np.tile(traj[:, -1:], (1, N))))                    # :264
    om = np.tile(OMEGA_DES, (orig.shape[1], 1)).T                              # :269-272
    trajectory = np.concatenate((orig[0:6, :], om))                            # :274-277
    second = np.gradient(np.gradient(trajectory[0:3, :], axis=1), axis=1) / dt ** 2   # :283
    nominal = np.vstack((second * MASS, np.zeros_like(second)))                # :285
    return trajectory, nominal


def hover_trajectory(duration, dt, position=(0.0, 0.0, 0.0)):
    """get_trajectory.generate_trajectory(form='point_stabilizing')  util/get_trajectory.py:71,109-124
    (t = arange(0, 10*duration, dt); identity quaternion [0,0,0,1]; zero rates)."""
    T = np.arange(0, 10 * duration, dt).size
    x = np.zeros((13, T))
    x[0], x[1], x[2] = position
    x[9] = 1.0
    return x


def window(trajectory, nominal, step, N):
    """get_next_trajectory_part with an integer step index (the reference uses int(t/dt), :360)."""
    return trajectory[:, step: step + N + 1].T.copy(), nominal[:, step: step + N + 1].T.copy()


def default_problem(N=15, step=0):
    """examples/sim.py default scenario at closed-loop step 0 (faults 10,11 stuck on, hover)."""
    from scipy.spatial.transform import Rotation
    fs = FaultSet([(10, 1.0), (11, 1.0)])
    x0 = np.concatenate([[1, 0, 1], [1, 0.5, 0],
                         Rotation.from_euler("zyx", [50, 30, -10], degrees=True).as_quat(),
                         [0.3, 0.8, -0.1]]).astype(float)                       # sim.py:49-54
    traj, nom = assign_trajectory(hover_trajectory(30, 0.1), N, 0.1)
    xr, ur = window(traj, nom, step, N)
    return Problem(fs, N, robot_to_center(x0), xr, ur), x0
