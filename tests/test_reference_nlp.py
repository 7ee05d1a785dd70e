"""Pinning the NLP, the allocator and the get_control pipeline against the REFERENCE'S OWN CONTROLLER CODE.

tests/golden/ref_nlp_fixtures.npz was produced by tools/gen_ref_nlp_fixtures.py, which executes (unmodified, from
/root/reference, in the build container) SpiralingController.__init__ -> set_model / set_cost_functions / build_solver
(spiraling_mpc.py:27-238), ControlAllocator.__init__ / get_physical_input (control_allocator.py:12-95) and
get_control (:288-317) on numeric stand-ins for casadi / cvxpy, and records
  * f(z, p), g(z, p), lbg, ubg of the reference's `nlp = dict(x, f, g, p)` at seeded decision vectors,
  * the allocation QP's data as the reference states it,
  * the thrust the reference's own post-processing + allocator return when the NLP solver hands back a given point.
The oracle's restatement (oracle.Problem.nlp_eval, oracle.allocate, oracle.get_control), the CPU port and -- with
-m gpu -- the CUDA path must reproduce them.  What remains unpinned by construction: WHICH local solution IPOPT
(tol 1e-3) would return; the fixtures prescribe the oracle's KKT point of the (now pinned) NLP.
"""
from pathlib import Path

import numpy as np
import pytest

from helpers import CpuPort, active_bits, host_tables

ROOT = Path(__file__).resolve().parent.parent
FIX = np.load(ROOT / "tests" / "golden" / "ref_nlp_fixtures.npz")
TAGS = [str(t) for t in FIX["tags"]]


def _faults(tag):
    return [(int(i), float(a)) for i, a in FIX[f"{tag}::faults"]]


def _problem(oracle, tag, c0):
    N = int(FIX[f"{tag}::N"])
    return oracle.Problem(oracle.FaultSet(_faults(tag)), N, np.asarray(c0, float), FIX[f"{tag}::x_ref"].T.copy(),
                          FIX[f"{tag}::u_ref"].T.copy())


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_nlp_equals_reference_build_solver(oracle, tag):
    """f, g (values AND row order), lbg, ubg of build_solver's nlp (spiraling_mpc.py:101-214) at 3 seeded z per scenario"""
    N = int(FIX[f"{tag}::N"])
    assert int(FIX[f"{tag}::num_var"]) == 6 * N + 13 * (N + 1)                                   # :110-115
    assert np.all(np.isneginf(FIX[f"{tag}::lbx"])) and np.all(np.isposinf(FIX[f"{tag}::ubx"]))   # :125-126
    for k in range(FIX[f"{tag}::z"].shape[0]):
        prob = _problem(oracle, tag, FIX[f"{tag}::x0"][k])
        f, g, lbg, ubg = prob.nlp_eval(FIX[f"{tag}::z"][k])
        assert np.isclose(f, FIX[f"{tag}::f"][k], rtol=1e-12, atol=0), (f, FIX[f"{tag}::f"][k])
        assert g.shape == FIX[f"{tag}::g"][k].shape == (13 * (N + 1) + prob.n_h * N + 72,)
        assert np.allclose(g, FIX[f"{tag}::g"][k], rtol=1e-12, atol=1e-12)
        assert np.array_equal(lbg, FIX[f"{tag}::lbg"])
        assert np.allclose(ubg, FIX[f"{tag}::ubg"], rtol=0, atol=1e-13)
        assert np.allclose(prob.u_comp, FIX[f"{tag}::u_comp"], atol=1e-15)                        # :141
    assert np.abs(FIX[f"{tag}::u_ref"]).max() > 0 or "hover" in tag                               # non-zero nominal wrench covered


def test_reference_solver_options():
    """ca.nlpsol('spiral_MPC_sol', 'ipopt', nlp, options)   spiraling_mpc.py:217-230"""
    for tag in TAGS:
        opts = dict(s.split("=", 1) for s in FIX[f"{tag}::options"])
        assert str(FIX[f"{tag}::plugin"]) == "ipopt"
        assert opts["ipopt.tol"] == "0.001" and opts["expand"] == "True" and opts["ipopt.print_level"] == "0"


@pytest.mark.parametrize("tag", TAGS)
def test_allocator_problem_equals_reference(oracle, tag):
    """ControlAllocator.__init__ (control_allocator.py:20-40): min sum_squares(u) s.t. u >= 0, u <= ub, D u == u_des"""
    fs = oracle.FaultSet(_faults(tag))
    assert str(FIX[f"{tag}::alloc_objective"]) == "min_sum_squares"
    assert np.array_equal(FIX[f"{tag}::alloc_D"], oracle.D_ALLOC)
    assert np.array_equal(FIX[f"{tag}::alloc_lb"], np.zeros(16)) and np.array_equal(FIX[f"{tag}::alloc_ub"], fs.ub)
    # the reference's own get_physical_input on the prescribed point: u_des and thrust
    thrust, ok = oracle.allocate(FIX[f"{tag}::pipe_udes"], FIX[f"{tag}::pipe_ub"])
    assert ok and np.allclose(thrust, FIX[f"{tag}::pipe_thrust"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_get_control_equals_reference_pipeline(oracle, tag):
    """get_control (:288-317): c0 = robot_to_center(x0), p = [c0; x_ref; u_ref] (column-major), u_res assembly, clip,
    allocation -- the reference's code ran with the NLP solver returning pipe_z"""
    N = int(FIX[f"{tag}::N"])
    c0 = oracle.robot_to_center(FIX[f"{tag}::pipe_robot"])
    assert np.allclose(c0, FIX[f"{tag}::pipe_c0"], rtol=1e-14, atol=1e-15)
    p = FIX[f"{tag}::pipe_p"]
    assert np.allclose(p[:13], c0, rtol=1e-14, atol=1e-15)                                       # :336
    assert np.array_equal(p[13:13 + 9 * (N + 1)], FIX[f"{tag}::x_ref"].T.ravel())                # order='F' flattening, :295
    assert np.array_equal(p[13 + 9 * (N + 1):], FIX[f"{tag}::u_ref"].T.ravel())                  # :296
    guess = FIX[f"{tag}::pipe_x0guess"]                                                          # cold start, :331-334
    assert np.all(guess[:6 * N] == 0) and np.allclose(guess[6 * N:6 * N + 13], c0, atol=1e-15) and np.all(guess[6 * N + 13:] == 0)
    z = FIX[f"{tag}::pipe_z"]
    warm = FIX[f"{tag}::pipe_x0guess_warm"]                                                      # warm start, :324-331
    assert np.array_equal(warm[:6 * (N - 1)], z[6:6 * N]) and np.all(warm[6 * (N - 1):6 * N] == 0)
    assert np.array_equal(warm[6 * N + 13:6 * N + 13 * N], z[6 * N + 26:]) and np.all(warm[6 * N + 13 * N:] == 0)
    prob = _problem(oracle, tag, c0)
    sol = {"U": z[:6 * N].reshape(N, 6), "X": z[6 * N:].reshape(N + 1, 13)}
    out = oracle.get_control(prob, sol)
    assert np.allclose(out["u_des"], FIX[f"{tag}::pipe_udes"], rtol=0, atol=1e-12)
    assert np.allclose(out["thrust"], FIX[f"{tag}::pipe_thrust"], rtol=0, atol=1e-9)
    # the prescribed point is a KKT point of the pinned NLP
    k = oracle.kkt_residual(prob, z[:6 * N])
    assert k["stat"] < 1e-8 and k["viol"] < 1e-9
    # reference quirk: the zero-tolerance membership test sends exact KKT points with an active hull
    # row into the undefined clip branch by <= 1 ulp -- the restated tolerance (clip_tol 1e-9) keeps them unchanged
    assert 0.0 <= float(FIX[f"{tag}::pipe_clip_excess"][0]) < 1e-15


def _run_case(tag, runner):
    N = int(FIX[f"{tag}::N"])
    cfg, table, masks, ffs, _ = host_tables([_faults(tag)], N)
    state = FIX[f"{tag}::pipe_robot"][None]
    xref = FIX[f"{tag}::x_ref"].T[None].copy()
    uref = FIX[f"{tag}::u_ref"].T[None].copy()
    return N, runner(cfg, table, state, xref, uref if np.abs(uref).max() > 0 else None, masks[:1], ffs[:1])


def _check_against_pipeline(oracle, tag, N, out):
    z = FIX[f"{tag}::pipe_z"]
    assert int(out["status"][0]) == 0
    u0 = z[:6]
    assert np.allclose(out["u0"][0], u0, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(u0).max()))
    assert np.allclose(out["thrust"][0], FIX[f"{tag}::pipe_thrust"], rtol=0, atol=2e-5)
    prob = _problem(oracle, tag, FIX[f"{tag}::pipe_c0"])
    c, _ = prob.ineq(z[:6 * N])
    want = [int(i) for i in np.where(c >= -oracle.ACTIVE_TOL)[0]]
    assert active_bits(out["active"][0], 26 * N + 72) == want


@pytest.mark.parametrize("tag", TAGS)
def test_cpu_port_equals_reference_pipeline(oracle, built, tag):
    port = CpuPort()
    N, out = _run_case(tag, lambda cfg, table, st, xr, ur, m, ff: port.step(cfg, table, st, xr, ur, m, ff, np.zeros(1, np.int32)))
    _check_against_pipeline(oracle, tag, N, out)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_equals_reference_pipeline(ft, oracle, built, tag):
    """the CUDA path (ftmpc_step through the C ABI) against the thrust the reference's get_control returned"""
    import torch
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
    from ft_mpc_b200.models import SystemModel
    N = int(FIX[f"{tag}::N"])
    eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, [_faults(tag)])
    d = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    uref = FIX[f"{tag}::u_ref"].T[None].copy()
    o = eng.step(d(FIX[f"{tag}::pipe_robot"][None]), d(FIX[f"{tag}::x_ref"].T[None].copy()),
                 d(uref) if np.abs(uref).max() > 0 else None)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in o.items() if k != "ws"}
    out["active"] = out["active"].view(np.uint32)
    _check_against_pipeline(oracle, tag, N, out)
