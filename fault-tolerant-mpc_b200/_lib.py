"""ctypes binding of libftmpc.so (include/ftmpc.h) and the host-side construction of ``ftmpc_config``.

Host logic only -- no numerics of the per-step path live here.  The constants mirror
ft_mpc/models/sys_model.py:52-131 and ft_mpc/controllers/tools/spiral_parameters.py:21-57 of the
reference; the terminal ingredients come from data/terminal.json (derived from the reference's
ft_mpc/config/terminal.yaml by tools/gen_terminal_data.py).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from pathlib import Path

import numpy as np

NX, NU, NE, NTHR, NH, NF = 13, 6, 9, 16, 26, 72
HULL_STRIDE = NH * NU + NH
MAX_POLY, MAX_ROOT = 32, 16

PHASE_NAMES = ['ls', 'lin', 'condense', 'cholesky', 'inverse', 'qp_setup', 'qp_active_set', 'qp_post', 'out',
               'ls_rollout', 'ls_eval', 'ls_terminal', 'chol_panel', 'chol_syrk', 'gi_select', 'gi_d', 'gi_z', 'gi_step', 'gi_update',
               'gi_drop', 'lin_jac', 'lin_costate', 'cond_pre', 'cond_col', 'cond_blk', 'warm_d0', 'warm_qr', 'warm_solve', 'warm_E',
               'ric_setup', 'ric_ab_step1', 'ric_ab_step2', 'ric_ab_wait', 'ric_cd', 'ric_cd_wait', 'ric_backsubst', 'ric_forward',
               'n_instances', 'n_sqp_iter', 'n_condense', 'n_chol_fail', 'n_qp', 'n_gi_iter', 'n_gi_drop', 'n_ls_backtrack', 'n_gi_warm_ok', 'n_gi_warm_miss', 'n_gi_refine']
N_PHASES = len(PHASE_NAMES)
N_CYCLE_PHASES = 37
ST_OK, ST_MAXITER, ST_QPFAIL, ST_INFEASIBLE, ST_ALLOC, ST_BADINPUT = 0, 1, 2, 3, 4, 5

_PKG = Path(__file__).resolve().parent
# FTMPC_LIB selects another build of the same library (A/B experiments on the GPU box); no fallback either way
LIB_PATH = Path(os.environ["FTMPC_LIB"]) if os.environ.get("FTMPC_LIB") else _PKG / "csrc" / "libftmpc.so"
DATA_DIR = _PKG / "data"


class FtmpcConfig(C.Structure):
    """struct ftmpc_config of include/ftmpc.h (field order and types must match exactly)."""
    _fields_ = [
        ("horizon", C.c_int32), ("dtype", C.c_int32), ("max_sqp_iter", C.c_int32), ("max_qp_iter", C.c_int32),
        ("stall_window", C.c_int32), ("warm_qp", C.c_int32), ("n_poly", C.c_int32), ("n_root", C.c_int32), ("n_hull_sets", C.c_int32),
        ("qp_method", C.c_int32),
        ("dt", C.c_double), ("mass", C.c_double), ("inertia", C.c_double * 3), ("r", C.c_double * 3),
        ("f_virt", C.c_double * 3), ("max_thrust", C.c_double),
        ("Q", C.c_double * NE), ("R", C.c_double * NU), ("D", C.c_double * (NU * NTHR)),
        ("Af", C.c_double * (NF * NE)), ("bf", C.c_double * NF),
        ("term_const", C.c_double),
        ("poly_c", C.c_double * MAX_POLY), ("poly_e", (C.c_int8 * NE) * MAX_POLY),
        ("root_c", C.c_double * MAX_ROOT), ("root_eps", C.c_double * MAX_ROOT), ("root_pow", C.c_double * MAX_ROOT),
        ("root_e", (C.c_int8 * NE) * MAX_ROOT),
        ("term_quad", C.c_double * (NE * NE)),
        ("sqp_tol", C.c_double), ("qp_tol", C.c_double), ("feas_tol", C.c_double), ("act_tol", C.c_double),
        ("rho_slack", C.c_double), ("clip_tol", C.c_double), ("theta_first", C.c_double), ("theta_growth", C.c_double), ("blend_dmax", C.c_double), ("fast_dmax", C.c_double),
    ]


def load_terminal(path=None) -> dict:
    """Terminal cost term table + terminal set (terminal_ingredients.py:451-474 equivalent, no eval)."""
    return json.loads(Path(path or DATA_DIR / "terminal.json").read_text())


def make_config(horizon: int, Q, R, *, dt: float, mass: float, inertia, r, f_virt, max_thrust: float, D,
                terminal: dict | None = None, n_hull_sets: int = 1, max_sqp_iter: int = 60, max_qp_iter: int = 0,
                stall_window: int = 10, sqp_tol: float = 1e-8, qp_tol: float = 1e-10, feas_tol: float = 1e-7,
                act_tol: float = 1e-7, rho_slack: float = 1e4, clip_tol: float = 1e-9, theta_first: float = 0.5,
                theta_growth: float = 2.0, blend_dmax: float = 1.0, warm_qp: int = 0, fast_dmax: float = 1e-5,
                qp_method: int = 0) -> FtmpcConfig:
    term = terminal or load_terminal()
    if len(term["poly"]) > MAX_POLY or len(term["root"]) > MAX_ROOT:
        raise ValueError("terminal cost has more terms than the term table holds")
    cfg = FtmpcConfig()
    cfg.horizon, cfg.dtype = int(horizon), 0
    cfg.max_sqp_iter = int(max_sqp_iter)
    n, m = NU * horizon, NH * horizon + NF + 2
    cfg.max_qp_iter = int(max_qp_iter) if max_qp_iter else 20 * (n + m)
    cfg.stall_window = int(stall_window)
    cfg.warm_qp = int(warm_qp)
    cfg.qp_method = int(qp_method)
    cfg.n_poly, cfg.n_root, cfg.n_hull_sets = len(term["poly"]), len(term["root"]), int(n_hull_sets)
    cfg.dt, cfg.mass, cfg.max_thrust = float(dt), float(mass), float(max_thrust)
    cfg.inertia[:] = [float(x) for x in np.diag(np.asarray(inertia, float))] if np.ndim(inertia) == 2 else list(map(float, inertia))
    cfg.r[:] = list(map(float, r))
    cfg.f_virt[:] = list(map(float, np.asarray(f_virt, float)[:3]))
    cfg.Q[:] = list(map(float, Q))
    cfg.R[:] = list(map(float, R))
    cfg.D[:] = list(map(float, np.asarray(D, float).reshape(-1)))
    A, b = np.asarray(term["A"], float), np.asarray(term["b"], float)
    assert A.shape == (NF, NE) and b.shape == (NF,)
    cfg.Af[:] = A.reshape(-1).tolist()
    cfg.bf[:] = b.tolist()
    cfg.term_const = float(term["const"])
    quad = np.zeros((NE, NE))
    for k, t in enumerate(term["poly"]):
        cfg.poly_c[k] = float(t["coeff"])
        for i, p in enumerate(t["exps"]):
            cfg.poly_e[k][i] = int(p)
        if sum(t["exps"]) == 2:          # pure quadratic part -> Gauss-Newton model of V_f
            idx = [i for i, p in enumerate(t["exps"]) for _ in range(p)]
            if idx[0] == idx[1]:
                quad[idx[0], idx[0]] += 2.0 * t["coeff"]
            else:
                quad[idx[0], idx[1]] += t["coeff"]
                quad[idx[1], idx[0]] += t["coeff"]
    if np.linalg.eigvalsh(quad).min() <= 0:
        raise ValueError("quadratic part of the terminal cost is not positive definite")
    cfg.term_quad[:] = quad.reshape(-1).tolist()
    for k, t in enumerate(term["root"]):
        cfg.root_c[k], cfg.root_eps[k], cfg.root_pow[k] = float(t["coeff"]), float(t["eps"]), float(t["pow"])
        for i, p in enumerate(t["exps"]):
            cfg.root_e[k][i] = int(p)
    cfg.sqp_tol, cfg.qp_tol, cfg.feas_tol = sqp_tol, qp_tol, feas_tol
    cfg.act_tol, cfg.rho_slack, cfg.clip_tol = act_tol, rho_slack, clip_tol
    cfg.theta_first, cfg.theta_growth, cfg.blend_dmax = theta_first, theta_growth, blend_dmax
    cfg.fast_dmax = fast_dmax
    return cfg


# keyword names of make_config that tune the solver (accepted in params["solver_opts"] / params["ftmpc_opts"])
SOLVER_OPTION_NAMES = ("max_sqp_iter", "max_qp_iter", "stall_window", "sqp_tol", "qp_tol", "feas_tol", "act_tol", "rho_slack",
                       "clip_tol", "theta_first", "theta_growth", "blend_dmax", "warm_qp", "fast_dmax", "qp_method")

_lib = None


def lib() -> C.CDLL:
    """Load libftmpc.so (built in-tree by __graft_entry__.build()).  Fails loudly -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  ft_mpc_b200 has no CPU fallback.")
    L = C.CDLL(os.fspath(LIB_PATH))
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
    L.ftmpc_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(FtmpcConfig), C.POINTER(C.c_double)]
    L.ftmpc_destroy.argtypes = [vp]
    L.ftmpc_destroy.restype = None
    L.ftmpc_strerror.argtypes = [C.c_int]
    L.ftmpc_strerror.restype = C.c_char_p
    L.ftmpc_workspace_bytes.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
    L.ftmpc_num_var.argtypes = [vp]
    L.ftmpc_num_ineq.argtypes = [vp]
    L.ftmpc_step.argtypes = [vp, C.c_int, dp, dp, dp, ip, dp, ip, C.c_int, dp, dp, dp, ip, ip, ip, dp, vp, C.c_size_t, vp]
    L.ftmpc_hull_facets.argtypes = [vp, C.c_int, ip, dp, dp, ip, ip, vp]
    L.ftmpc_rk4_jac.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, vp]
    L.ftmpc_robot_to_center.argtypes = [vp, C.c_int, dp, dp, vp]
    L.ftmpc_terminal.argtypes = [vp, C.c_int, dp, dp, dp, dp, vp]
    L.ftmpc_condense.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp, dp, C.c_double, dp, dp, vp]
    L.ftmpc_qp_solve.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, ip, ip, dp, dp, dp, dp, ip, vp]
    L.ftmpc_allocate.argtypes = [vp, C.c_int, dp, dp, dp, ip, vp]
    L.ftmpc_closed_loop.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, ip, dp, ip, dp, C.c_int, dp, dp, dp, ip, ip,
                                    ip, dp, dp, ip, vp, C.c_size_t, vp]
    L.ftmpc_closed_loop.restype = C.c_int
    L.ftmpc_clip.argtypes = [vp, C.c_int, ip, dp, dp, ip, vp]
    L.ftmpc_clip.restype = C.c_int
    L.ftmpc_plant_step.argtypes = [vp, C.c_int, dp, dp, ip, dp, dp, C.c_int, dp, vp]
    L.ftmpc_profile_enable.argtypes = [vp, C.c_int]
    L.ftmpc_profile_read.argtypes = [vp, vp, dp, ip, C.c_int]
    L.ftmpc_last_launches.argtypes = [vp]
    L.ftmpc_fp64_peak.argtypes = [vp, C.POINTER(C.c_double), vp]
    L.ftmpc_fp64_peak.restype = C.c_int
    for name in ("ftmpc_profile_enable", "ftmpc_profile_read", "ftmpc_last_launches", "ftmpc_create", "ftmpc_workspace_bytes", "ftmpc_num_var", "ftmpc_num_ineq", "ftmpc_step",
                 "ftmpc_rk4_jac", "ftmpc_robot_to_center", "ftmpc_terminal", "ftmpc_condense", "ftmpc_qp_solve",
                 "ftmpc_allocate", "ftmpc_plant_step"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(code: int, what: str = "ftmpc") -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed: {lib().ftmpc_strerror(code).decode()} ({code})")
