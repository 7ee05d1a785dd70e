"""CPU tests of the solver ALGORITHM: the product's per-instance headers compiled for the host
(oracle/cpu_port, SerialBlock instantiation -- test infrastructure, never loaded by the product) against
the independent oracle and the golden vectors.  The same checks run against the CUDA library in
test_gpu_parity.py; this file lets the algorithm be verified without a GPU."""
import ctypes as C

import numpy as np
import pytest

import helpers as H


@pytest.fixture(scope="module")
def port(built):
    return H.CpuPort()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_rk4_jacobians_vs_complex_step(port, oracle):
    """K1: x_{t+1}, [A_t | B_t] columns in z-order [w q F tau] vs the oracle's complex-step derivatives"""
    N = 6
    cfg, *_ = H.host_tables([[(3, 0.0)]], N)
    rng = np.random.default_rng(0)
    B = 3
    x = np.zeros((B, N + 1, 13)); x[:, 0] = rng.normal(0, 0.7, (B, 13))
    W = rng.normal(0, 2.0, (B, N, 6))
    jac = np.zeros((B, N, 13, 13))
    port.lib.ftmpc_cpu_rk4_jac(C.byref(cfg), B, _p(x), _p(W), _p(jac), None, None)
    df = np.zeros(6)
    for b in range(B):
        for t in range(N):
            xn = oracle.spiral_rk4(x[b, t], W[b, t], df, 0.1)
            assert np.allclose(x[b, t + 1], xn, rtol=1e-13, atol=1e-13)
            h = 1e-30
            for c in range(13):
                xp, wp = x[b, t].astype(complex), W[b, t].astype(complex)
                if c < 7:
                    xp[6 + c] += 1j * h
                else:
                    wp[c - 7] += 1j * h
                col = oracle.spiral_rk4(xp, wp, df, 0.1).imag / h
                assert np.allclose(jac[b, t, c], col, rtol=1e-11, atol=1e-12), (b, t, c)


def test_terminal_cost_grad_hess(port, oracle):
    cfg, *_ = H.host_tables([[(3, 0.0)]], 5)
    rng = np.random.default_rng(1)
    E = np.vstack([rng.normal(0, 0.2, (6, 9)), [a["e"] for a in oracle.TERMINAL.anchors[1:]]])
    B = E.shape[0]
    V, g, Hs = np.zeros(B), np.zeros((B, 9)), np.zeros((B, 81))
    port.lib.ftmpc_cpu_terminal(C.byref(cfg), B, _p(E), _p(V), _p(g), _p(Hs))
    for b in range(B):
        assert V[b] == pytest.approx(float(oracle.TERMINAL.cost(E[b])), rel=1e-12, abs=1e-10)
        assert np.allclose(g[b], oracle.TERMINAL.grad(E[b]), rtol=1e-10, atol=1e-9)
        Hb = Hs[b].reshape(9, 9)
        assert np.allclose(Hb, Hb.T, atol=1e-12)
        for i in range(9):
            d = np.zeros(9); d[i] = 1e-6
            fd = (oracle.TERMINAL.grad(E[b] + d) - oracle.TERMINAL.grad(E[b] - d)) / 2e-6
            assert np.allclose(Hb[i], fd, rtol=1e-5, atol=1e-4)
    assert V[-2] == pytest.approx(82.584936488841, rel=1e-12)          # SURVEY 8c anchors (from terminal.yaml)
    assert V[-1] == pytest.approx(40.774178314212, rel=1e-12)


def test_allocator_vs_oracle(port, oracle):
    cfg, *_ = H.host_tables([[(10, 1.0), (11, 1.0)]], 5)
    rng = np.random.default_rng(2)
    ub = np.full(16, 3.4); ub[[10, 11]] = 0.0
    B = 12
    udes = np.stack([oracle.D_ALLOC @ (rng.uniform(0, 3.4, 16) * (ub > 0)) for _ in range(B)])
    ubs = np.tile(ub, (B, 1))
    th, st = np.zeros((B, 16)), np.zeros(B, np.int32)
    port.lib.ftmpc_cpu_allocate(C.byref(cfg), B, _p(udes), _p(ubs), _p(th), _p(st))
    assert (st == 0).all()
    for b in range(B):
        tho, ok = oracle.allocate(udes[b], ub)
        assert ok and np.allclose(th[b], tho, atol=1e-7)
        assert np.allclose(oracle.D_ALLOC @ th[b], udes[b], atol=1e-9)
    # infeasible request -> status != 0 (control_allocator.py:88-93 exit()s; the library reports)
    bad = np.array([[100.0, 0, 0, 0, 0, 0]])
    port.lib.ftmpc_cpu_allocate(C.byref(cfg), 1, _p(bad), _p(ubs[:1].copy()), _p(th[:1].copy()), _p(st[:1]))
    assert st[0] != 0


def test_generic_qp_kkt(port):
    """K3 algorithm: dual active-set QP on random strictly convex problems with sparse rows; KKT conditions"""
    rng = np.random.default_rng(4)
    n, m = 30, 50
    for trial in range(4):
        M = rng.normal(size=(n, n)); G = M @ M.T + n * np.eye(n); a = rng.normal(0, 5, n)
        ptr, idx, val = [0], [], []
        for i in range(m):
            cols = rng.choice(n, 4, replace=False)
            idx += list(cols); val += list(rng.normal(size=4)); ptr.append(len(idx))
        ptr, idx, val = np.array(ptr, np.int32), np.array(idx, np.int32), np.array(val)
        beta = -rng.uniform(0.0, 0.3, m)                     # n_i'x >= beta_i, x = 0 feasible
        x, lam = np.zeros(n), np.zeros(m)
        it, na = C.c_int(), C.c_int()
        st = port.lib.gi_test(n, m, 0, _p(G), _p(a), _p(ptr), _p(idx), _p(val), _p(beta), _p(x), _p(lam), C.byref(it), C.byref(na))
        assert st == 0
        Cn = np.zeros((m, n))
        for i in range(m):
            Cn[i, idx[ptr[i]:ptr[i + 1]]] = val[ptr[i]:ptr[i + 1]]
        s = Cn @ x - beta
        assert s.min() > -1e-9 and lam.min() >= 0
        assert np.abs(G @ x + a - Cn.T @ lam).max() < 1e-8            # stationarity
        assert np.abs(lam * s).max() < 1e-8                           # complementarity


@pytest.mark.parametrize("N", [15, 20])
def test_full_step_vs_golden(port, oracle, golden, N):
    """whole get_control (SQP -> u0 -> allocation) on the golden instances: same KKT point as the oracle"""
    ks = [k for k in H.cases_with_horizon(golden, N) if not golden["warm"][k]]
    sets, scen = H.gather_cases(golden, ks)
    cfg, table, masks, ffs, _ = H.host_tables(sets, N)
    out = port.step(cfg, table, golden["x0"][ks], golden["xref"][ks][:, :N + 1], golden["uref"][ks][:, :N + 1],
                    masks[scen], ffs[scen], scen)
    assert (out["status"] == 0).all(), out["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        assert out["cost"][j] == pytest.approx(golden["f"][k], rel=1e-9), golden["name"][k]
        u0 = golden["U"][k, 0]
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), golden["name"][k]
        assert np.allclose(out["z"][j, :6 * N], golden["U"][k, :N].ravel(), atol=2e-5)
        assert H.active_bits(out["active"][j], nbits) == H.active_bits(golden["active"][k], nbits), golden["name"][k]
        assert np.allclose(out["thrust"][j], golden["thrust"][k], atol=2e-5), golden["name"][k]


@pytest.mark.parametrize("N", [15, 20])
def test_accelerating_reference_vs_golden(port, accel_golden, N):
    """row f-3: circle references, nominal wrench rotated into the body frame per stage (spiraling_mpc.py:156-166)"""
    g = accel_golden
    ks = H.cases_with_horizon(g, N)
    assert len(ks) == 2 and np.abs(g["uref"][ks]).max() > 1.0
    sets, scen = H.gather_cases(g, ks)
    cfg, table, masks, ffs, _ = H.host_tables(sets, N)
    out = port.step(cfg, table, g["x0"][ks], g["xref"][ks][:, :N + 1], g["uref"][ks][:, :N + 1], masks[scen], ffs[scen], scen)
    assert (out["status"] == 0).all(), out["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        u0 = g["U"][k, 0]
        assert out["cost"][j] == pytest.approx(g["f"][k], rel=1e-9), g["name"][k]
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), g["name"][k]
        assert H.active_bits(out["active"][j], nbits) == H.active_bits(g["active"][k], nbits), g["name"][k]
        assert np.allclose(out["thrust"][j], g["thrust"][k], atol=2e-5), g["name"][k]


def test_bench_workload_instances_vs_golden(port, bench_golden):
    """32 instances of the bench workload (single / double faults, dead / stuck-on): the solver that bench.py times lands on
    the oracle's KKT point for every one of them"""
    g = bench_golden
    N = 20
    ks = H.cases_with_horizon(g, N)
    assert len(ks) == 32
    sets, scen = H.gather_cases(g, ks)
    cfg, table, masks, ffs, _ = H.host_tables(sets, N)
    out = port.step(cfg, table, g["x0"][ks], g["xref"][ks][:, :N + 1], None, masks[scen], ffs[scen], scen)
    assert (out["status"] == 0).all(), out["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        u0 = g["U"][k, 0]
        assert out["cost"][j] == pytest.approx(g["f"][k], rel=1e-9), g["name"][k]
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), g["name"][k]
        assert H.active_bits(out["active"][j], nbits) == H.active_bits(g["active"][k], nbits), g["name"][k]
        assert np.allclose(out["thrust"][j], g["thrust"][k], atol=2e-5), g["name"][k]


def test_warm_start_closed_loop_vs_golden(port, oracle, golden):
    """warm start = previous solution shifted one stage (spiraling_mpc.py:324-331); closed-loop steps 1..4"""
    N = 15
    ks = [k for k in range(len(golden["N"])) if golden["warm"][k]]
    assert len(ks) == 4
    cfg, table, masks, ffs, _ = H.host_tables([H.case_faults(golden, ks[0])], N)
    z = np.stack([H.z_from_U0(golden["U0"][k, :N], N) for k in ks])
    out = port.step(cfg, table, golden["x0"][ks], golden["xref"][ks][:, :N + 1], golden["uref"][ks][:, :N + 1],
                    masks[[0] * 4], ffs[[0] * 4], np.zeros(4, np.int32), warm=1, z=z)
    assert (out["status"] == 0).all()
    cold = port.step(cfg, table, golden["x0"][ks], golden["xref"][ks][:, :N + 1], golden["uref"][ks][:, :N + 1],
                     masks[[0] * 4], ffs[[0] * 4], np.zeros(4, np.int32))
    for j, k in enumerate(ks):
        assert np.abs(out["u0"][j] - golden["U"][k, 0]).max() <= 1e-5 * max(1.0, np.abs(golden["U"][k, 0]).max())
        assert np.allclose(out["thrust"][j], golden["thrust"][k], atol=2e-5)
    assert out["iters"][:, 0].sum() < cold["iters"][:, 0].sum()       # the warm start pays off


def test_infeasible_instance_is_flagged(port, golden):
    k = list(golden["name"]).index("line_N15")
    N = 15
    cfg, table, masks, ffs, _ = H.host_tables([H.case_faults(golden, k)], N)
    out = port.step(cfg, table, golden["x0"][[k]], golden["xref"][[k]][:, :N + 1], golden["uref"][[k]][:, :N + 1],
                    masks[[0]], ffs[[0]], np.zeros(1, np.int32))
    assert out["status"][0] != 0
    assert np.isfinite(out["thrust"]).all() and (out["thrust"] >= 0).all() and (out["thrust"] <= 3.4 + 1e-12).all()


def _bench_sample(port, n, offset=0, **opts):
    """first n instances (from `offset`) of bench.py's workload through the CPU checker with ftmpc_config overrides"""
    import bench
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200 import _lib as L
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, hull_table_entry
    from ft_mpc_b200.controllers.tools.spiral_parameters import SpiralParameters
    from ft_mpc_b200.models import SystemModel
    cells, states, scen, xref = bench.make_workload(8192, 20, 1)        # the bench batch (instance numbers refer to it)
    model = SystemModel(0.1)
    sp = SpiralParameters(model)
    table = np.ascontiguousarray(np.stack([hull_table_entry(c["A"], c["b"]) for c in cells]))
    cfg = L.make_config(20, DEFAULT_Q, DEFAULT_R, dt=model.dt, mass=model.mass, inertia=model.inertia, r=sp.r, f_virt=sp.f_virt,
                        max_thrust=model.max_thrust, D=model.D, n_hull_sets=len(cells), **opts)
    masks = np.zeros(len(cells), np.uint16)
    ffs = np.zeros((len(cells), 16))
    for k, c in enumerate(cells):
        for i, a in c["faults"]:
            masks[k] |= np.uint16(1 << i)
            ffs[k, i] = a * model.max_thrust
    s = slice(offset, offset + n)
    return port.step(cfg, table, states[s], xref[s], None, masks[scen[s]], ffs[scen[s]], scen[s])


def test_early_stop_in_the_quadratic_regime_changes_nothing_but_the_count(port):
    """ftmpc_config.fast_dmax: skipping the QP that only confirms convergence saves iterations and moves the answer by less
    than the termination tolerance (u0 1e-8 relative here; the parity bar is 1e-5), active sets unchanged"""
    fast = _bench_sample(port, 96)
    full = _bench_sample(port, 96, fast_dmax=0.0)
    ok = (fast["status"] == 0) & (full["status"] == 0)
    assert ok.sum() >= 95 and np.array_equal(fast["status"] == 0, full["status"] == 0)
    assert fast["iters"][ok, 0].sum() < full["iters"][ok, 0].sum()
    assert (fast["iters"][:, 0] <= full["iters"][:, 0]).all()
    du = np.abs(fast["u0"] - full["u0"]).max(axis=1) / np.maximum(1.0, np.abs(full["u0"]).max(axis=1))
    assert du[ok].max() < 1e-8
    assert np.array_equal(fast["active"][ok], full["active"][ok])
    assert np.abs(fast["thrust"] - full["thrust"])[ok].max() < 1e-7


def test_stall_detector_only_stops_hopeless_instances(port):
    """ftmpc_config.stall_window: instance 249 of the bench workload creeps along a flat non-convex valley (step ~2e-4,
    objective moving by 1e-11 relative per iteration; it reaches a KKT point after 250 iterations, a different local minimum
    than the oracle's -- tools/hard_instances.py) and is given up after 30 iterations instead of the cap of 60; every instance
    that converges without the detector still converges with it, to the same point"""
    on = _bench_sample(port, 64, offset=224)
    off = _bench_sample(port, 64, offset=224, stall_window=0)
    k = 249 - 224
    assert off["status"][k] == 1 and off["iters"][k, 0] == 60
    assert on["status"][k] == 1 and on["iters"][k, 0] <= 30
    conv = off["status"] == 0
    assert conv.sum() == 63 and np.array_equal(on["status"] == 0, conv)
    assert np.array_equal(on["iters"][conv], off["iters"][conv]) and np.array_equal(on["u0"][conv], off["u0"][conv])
