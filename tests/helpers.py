"""Shared test helpers: golden-case decoding, the CPU checker binding, KKT checks through the oracle."""
import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
CPU_LIB = ROOT / "oracle" / "_cpu" / "libftmpc_cpu.so"


def case_faults(g, k):
    return [(int(i), float(a)) for i, a in zip(g["fault_idx"][k], g["fault_inten"][k]) if i >= 0]


def case_problem(o, g, k):
    """oracle.Problem of golden case k"""
    N = int(g["N"][k])
    return o.Problem(o.FaultSet(case_faults(g, k)), N, o.robot_to_center(g["x0"][k]), g["xref"][k, :N + 1].copy(),
                     g["uref"][k, :N + 1].copy())


def feasible_cases(g):
    """cases whose NLP the oracle solved to a KKT point (one golden case, line_N15, is an infeasible NLP:
    the terminal set cannot be reached; the library must flag it with status != 0)"""
    return np.asarray(g["kkt_viol"]) < 1e-8


def cases_with_horizon(g, N, feasible_only=True):
    ok = feasible_cases(g)
    return [k for k in range(len(g["N"])) if int(g["N"][k]) == N and (ok[k] or not feasible_only)]


def active_bits(words, nbits):
    words = np.asarray(words).astype(np.uint32)
    return [i for i in range(nbits) if (int(words[i // 32]) >> (i % 32)) & 1]


def z_from_U0(U0, N):
    """decision vector [u | x] whose shifted control block reproduces U0 under the warm start (:324-331):
    the library takes u_t <- z[u_{t+1}], so store U0 one stage later."""
    z = np.zeros(6 * N + 13 * (N + 1))
    z[6:6 * N] = np.asarray(U0).reshape(N, 6)[:N - 1].ravel()
    return z


class CpuPort:
    """ctypes binding of oracle/_cpu/libftmpc_cpu.so (the product headers compiled with SerialBlock)."""

    def __init__(self):
        self.lib = C.CDLL(str(CPU_LIB))
        self.lib.ftmpc_cpu_num_threads.restype = C.c_int

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None

    def step(self, cfg, hull_table, state, xref, uref, mask, ff, hidx, warm=0, z=None, nthreads=0):
        B = state.shape[0]
        N = cfg.horizon
        nz = 6 * N + 13 * (N + 1)
        mc = 26 * N + 72
        z = np.zeros((B, nz)) if z is None else np.ascontiguousarray(z, dtype=float)
        out = dict(z=z, thrust=np.zeros((B, 16)), u0=np.zeros((B, 6)), active=np.zeros((B, (mc + 31) // 32), np.uint32),
                   status=np.zeros(B, np.int32), iters=np.zeros((B, 2), np.int32), cost=np.zeros(B))
        p = self._p
        self.lib.ftmpc_cpu_step(C.byref(cfg), p(np.ascontiguousarray(hull_table)), B, p(np.ascontiguousarray(state)),
                                p(np.ascontiguousarray(xref)), p(np.ascontiguousarray(uref)) if uref is not None else None,
                                p(np.ascontiguousarray(mask, dtype=np.uint16)), p(np.ascontiguousarray(ff)),
                                p(np.ascontiguousarray(hidx, dtype=np.int32)), int(warm), p(z), p(out["thrust"]), p(out["u0"]),
                                p(out["active"]), p(out["status"]), p(out["iters"]), p(out["cost"]), None, int(nthreads))
        return out


def host_tables(fault_sets, N, Q=None, R=None, **opts):
    """ftmpc_config + hull table + per-set (mask, fault_force) built by the product's host code (no CUDA needed)."""
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200 import _lib as L
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, hull_table_entry
    from ft_mpc_b200.controllers.tools.input_bounds import hull_of_faults
    from ft_mpc_b200.controllers.tools.spiral_parameters import SpiralParameters
    from ft_mpc_b200.models import SystemModel
    model = SystemModel(0.1)
    sp = SpiralParameters(model)
    table, masks, ffs = [], [], []
    for fs in fault_sets:
        A, b = hull_of_faults(model.D, model.max_thrust, fs)
        table.append(hull_table_entry(A, b))
        ff = np.zeros(16)
        m = 0
        for i, a in fs:
            ff[i] = a * model.max_thrust
            m |= 1 << i
        masks.append(m)
        ffs.append(ff)
    cfg = L.make_config(N, Q or DEFAULT_Q, R or DEFAULT_R, dt=model.dt, mass=model.mass, inertia=model.inertia, r=sp.r,
                        f_virt=sp.f_virt, max_thrust=model.max_thrust, D=model.D, n_hull_sets=len(fault_sets), **opts)
    return cfg, np.ascontiguousarray(np.stack(table)), np.array(masks, np.uint16), np.stack(ffs), model


def gather_cases(g, ks):
    """distinct fault sets + per-case scenario index for golden cases ks (all with the same horizon)"""
    sets, scen = [], []
    for k in ks:
        fs = case_faults(g, k)
        if fs not in sets:
            sets.append(fs)
        scen.append(sets.index(fs))
    return sets, np.array(scen)
