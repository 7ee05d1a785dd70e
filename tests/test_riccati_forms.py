"""The three forms of the QP sub-problem must land on the same KKT point -- the oracle's:

  * default at N <= 20: Riccati factorisation J = Phi^-1 blkdiag(C_t^-T) (csrc/ftmpc_riccati.cuh) + null-space active set;
  * qp_method bit 4 (host builds only): the round-1 path, condensed Hessian + Cholesky + explicit L^-T;
  * operator form (default at N > 20, forced by qp_method bit 5, forbidden by bit 6): range-space active set with
    K = E E' applied through the stage records (ric_apply / ric_apply_g), nothing of size N^2 in memory.

CPU: the product headers compiled for the host (oracle/cpu_port) against the goldens (the oracle's KKT points).
GPU (-m gpu): the CUDA library on the same instances.
"""
import numpy as np
import pytest

import helpers as H

FORMS = {"riccati+null-space": 0, "dense factor": 16, "operator two-interval": 32}


@pytest.fixture(scope="module")
def port(built):
    return H.CpuPort()


def _solve_cpu(port, g, ks, N, qp_method):
    sets, scen = H.gather_cases(g, ks)
    cfg, table, masks, ffs, _ = H.host_tables(sets, N, qp_method=qp_method)
    return port.step(cfg, table, g["x0"][ks], g["xref"][ks][:, :N + 1], g["uref"][ks][:, :N + 1], masks[scen], ffs[scen], scen)


@pytest.mark.parametrize("form", list(FORMS))
def test_cpu_forms_vs_golden(port, golden, form):
    N = 20
    ks = [k for k in H.cases_with_horizon(golden, N) if not golden["warm"][k]]
    out = _solve_cpu(port, golden, ks, N, FORMS[form])
    assert (out["status"] == 0).all(), out["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        u0 = golden["U"][k, 0]
        assert out["cost"][j] == pytest.approx(golden["f"][k], rel=1e-9), (form, golden["name"][k])
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), (form, golden["name"][k])
        assert H.active_bits(out["active"][j], nbits) == H.active_bits(golden["active"][k], nbits), (form, golden["name"][k])


def test_cpu_forms_take_the_same_iterations(port, bench_golden):
    """same QP, same pivoting rules: the three forms differ by rounding only, so the iteration counts agree (up to the last,
    confirming iteration)"""
    g, N = bench_golden, 20
    ks = H.cases_with_horizon(g, N)[:12]
    ref = _solve_cpu(port, g, ks, N, 0)
    for form, qm in FORMS.items():
        out = _solve_cpu(port, g, ks, N, qm)
        assert (out["status"] == ref["status"]).all(), form
        # (an iterate sitting on the termination threshold may take one confirming iteration more or less)
        assert np.abs(out["iters"][:, 0] - ref["iters"][:, 0]).max() <= 1, (form, out["iters"][:, 0], ref["iters"][:, 0])
        assert (out["iters"][:, 0] == ref["iters"][:, 0]).mean() >= 0.9, (form, out["iters"][:, 0], ref["iters"][:, 0])
        assert np.abs(out["u0"] - ref["u0"]).max() < 2e-6, form      # (the dense path scales the convexification by max diag H, the Riccati path by max diag Lam_t: same KKT point, last step differs)


def _long_inputs(N, B, seed):
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200.util import scenarios
    cells = scenarios.load_cells(kinds=("single",))[:4]
    st = scenarios.random_states(B, seed)
    scen = np.arange(B) % len(cells)
    return cells, st, scen, scenarios.hover_reference(B, N)


@pytest.mark.parametrize("N", [30, 60])
def test_cpu_long_horizon_operator_g_form_equals_dense(port, N):
    """N > 20: the default is the operator form on the stage matrices G_t (ric_apply_g); bit 6 keeps the dense null-space form"""
    B = 4
    cells, st, scen, xref = _long_inputs(N, B, 21)
    sets = [c["faults"] for c in cells]
    outs = {}
    for qm in (0, 64):
        cfg, table, masks, ffs, _ = H.host_tables(sets, N, qp_method=qm)
        outs[qm] = port.step(cfg, table, st, xref, None, masks[scen], ffs[scen], scen)
    a, b = outs[0], outs[64]
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.abs(a["iters"][:, 0] - b["iters"][:, 0]).max() <= 1
    assert np.abs(a["u0"] - b["u0"]).max() < 2e-6
    assert np.allclose(a["cost"], b["cost"], rtol=1e-10)
    nbits = 26 * N + 72
    for j in range(B):
        assert H.active_bits(a["active"][j], nbits) == H.active_bits(b["active"][j], nbits)


@pytest.mark.gpu
@pytest.mark.parametrize("qp_method", [0, 32])
def test_gpu_forms_vs_golden(ft, built, golden, qp_method):
    """N = 20 on the device: the default (Riccati + null-space) and the forced operator form (one warp sweeps the stage
    records in shared memory) against the oracle's KKT points"""
    import torch
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
    from ft_mpc_b200.models import SystemModel
    N = 20
    ks = [k for k in H.cases_with_horizon(golden, N) if not golden["warm"][k]]
    sets, scen = H.gather_cases(golden, ks)
    eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, sets, qp_method=qp_method)
    d = lambda a, t=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=t, device="cuda")
    out = eng.step(d(golden["x0"][ks]), d(golden["xref"][ks][:, :N + 1]), scenario=d(scen, torch.int64))
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in out.items() if k != "ws"}
    assert (g["status"] == 0).all(), g["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        u0 = golden["U"][k, 0]
        assert np.abs(g["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), golden["name"][k]
        assert H.active_bits(g["active"][j].view(np.uint32), nbits) == H.active_bits(golden["active"][k], nbits), golden["name"][k]


@pytest.mark.gpu
def test_gpu_long_horizon_equals_cpu_port(ft, built):
    """N = 30: the staged G-form operator on the device against the host build of the same headers (same iterations, u0 1e-7)"""
    import torch
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
    from ft_mpc_b200.models import SystemModel
    N, B = 30, 6
    cells, st, scen, xref = _long_inputs(N, B, 33)
    eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, cells)
    d = lambda a, t=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=t, device="cuda")
    out = eng.step(d(st), d(xref), scenario=d(scen, torch.int64))
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in out.items() if k != "ws"}
    cfg, table, masks, ffs, _ = H.host_tables([c["faults"] for c in cells], N)
    ref = H.CpuPort().step(cfg, table, st, xref, None, masks[scen], ffs[scen], scen)
    assert (g["status"] == ref["status"]).all()
    assert np.abs(g["iters"][:, 0] - ref["iters"][:, 0]).max() <= 1
    assert np.abs(g["u0"] - ref["u0"]).max() < 2e-6



def test_cpu_long_horizon_accelerating_reference_operator_equals_dense(port):
    """N = 30 with a non-zero nominal wrench (the input-cost coupling terms of ftmpc_riccati.cuh on the operator path)"""
    N, B = 30, 3
    cells, st, scen, xref = _long_inputs(N, B, 7)
    uref = np.zeros((B, N + 1, 6)); uref[:, :, 0] = 1.5; uref[:, :, 1] = -0.7; uref[:, :, 2] = 0.4
    sets = [c["faults"] for c in cells]
    outs = {}
    for qm in (0, 64):
        cfg, table, masks, ffs, _ = H.host_tables(sets, N, qp_method=qm)
        outs[qm] = port.step(cfg, table, st, xref, uref, masks[scen], ffs[scen], scen)
    a, b = outs[0], outs[64]
    assert (a["status"] == b["status"]).all() and (a["status"] == 0).any()
    ok = a["status"] == 0
    assert np.abs(a["iters"][ok, 0] - b["iters"][ok, 0]).max() <= 1
    assert np.abs(a["u0"][ok] - b["u0"][ok]).max() < 2e-6
    assert np.allclose(a["cost"][ok], b["cost"][ok], rtol=1e-9)
