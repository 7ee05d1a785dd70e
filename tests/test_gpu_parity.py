"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI of libftmpc.so
(include/ftmpc.h) -- directly for the stage entry points, through the host mirror (BatchedMPC /
SpiralingController) for ftmpc_step -- and is compared with the CPU oracle (oracle/ftmpc_oracle.py), the
committed golden vectors (tests/golden/nlp_cases.npz) and size-independent properties at full batch sizes.

Tolerances (BASELINE.json north_star, fp64): first control input within 1e-5 relative, fault masks and
active-set indices bit-exact, thrust within 2e-5 absolute, closed-loop state within 1e-6."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
F64 = torch.float64


def dev(a, dtype=F64):
    return torch.tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


_KEEP = []       # tensors whose device pointers were handed to the C ABI stay alive until the test module ends


def ptr(t):
    """raw device pointer for the C ABI; the tensor is kept alive (a temporary passed as ptr(dev(x)) would be
    returned to the caching allocator -- and reused by the next temporary -- before the kernel runs)"""
    if t is None:
        return None
    _KEEP.append(t)
    if len(_KEEP) > 4096:
        torch.cuda.synchronize()
        del _KEEP[:2048]
    return C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def L(ft, built):
    from ft_mpc_b200 import _lib
    _lib.lib()
    return _lib


def make_engine(fault_sets, N, **opts):
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
    from ft_mpc_b200.models import SystemModel
    return BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, fault_sets, **opts)


@pytest.fixture(scope="module")
def eng6(L):
    return make_engine([[(3, 0.0)]], 6)


# ---------------------------------------------------------------------------------------------------------
# stage entry points
# ---------------------------------------------------------------------------------------------------------
def test_rk4_jac_vs_oracle(L, eng6, oracle):
    N, B = 6, 37                                    # ragged: not a multiple of the 4 warps per block
    rng = np.random.default_rng(0)
    x = np.zeros((B, N + 1, 13)); x[:, 0] = rng.normal(0, 0.7, (B, 13))
    W = rng.normal(0, 2.0, (B, N, 6))
    lam = rng.normal(0, 1.0, (B, N + 1, 13))
    xd, Wd, lamd = dev(x), dev(W), dev(lam)
    jac = torch.zeros(B, N, 13, 13, dtype=F64, device="cuda")
    hess = torch.zeros(B, N, 13, 13, dtype=F64, device="cuda")
    L.check(L.lib().ftmpc_rk4_jac(eng6.handle, B, ptr(xd), ptr(Wd), ptr(jac), ptr(lamd), ptr(hess), None))
    torch.cuda.synchronize()
    xg, jg, hg = xd.cpu().numpy(), jac.cpu().numpy(), hess.cpu().numpy()
    df, h = np.zeros(6), 1e-30

    def zmap(z, xb):        # z-space [w q F tau] -> next state
        xx = np.concatenate([xb[:6], z[:7]])
        return oracle.spiral_rk4(xx, z[7:], df, 0.1)

    for b in (0, 17, 36):
        for t in range(N):
            assert np.allclose(xg[b, t + 1], oracle.spiral_rk4(xg[b, t], W[b, t], df, 0.1), rtol=1e-13, atol=1e-13)
            z0 = np.concatenate([xg[b, t, 6:13], W[b, t]])
            J = np.zeros((13, 13))
            for c in range(13):
                zp = z0.astype(complex); zp[c] += 1j * h
                J[c] = zmap(zp, xg[b, t].astype(complex)).imag / h
            assert np.allclose(jg[b, t], J, rtol=1e-11, atol=1e-12)
            # Hessian of lam_{t+1}' RK4 in z-space: central differences of the complex-step gradient
            def grad(z):
                g = np.zeros(13)
                for c in range(13):
                    zp = z.astype(complex); zp[c] += 1j * h
                    g[c] = lam[b, t + 1] @ zmap(zp, xg[b, t].astype(complex)).imag / h
                return g
            Hfd = np.zeros((13, 13))
            for c in range(13):
                d = np.zeros(13); d[c] = 1e-5
                Hfd[c] = (grad(z0 + d) - grad(z0 - d)) / 2e-5
            assert np.allclose(hg[b, t], Hfd, rtol=1e-6, atol=1e-6), (b, t)


def test_robot_to_center_vs_oracle(L, eng6, oracle):
    from ft_mpc_b200.util import scenarios
    st = scenarios.random_states(130, 3)
    c = torch.empty(130, 13, dtype=F64, device="cuda")
    L.check(L.lib().ftmpc_robot_to_center(eng6.handle, 130, ptr(dev(st)), ptr(c), None))
    assert np.allclose(c.cpu().numpy(), oracle.robot_to_center(st), rtol=1e-14, atol=1e-14)


def test_terminal_vs_oracle(L, eng6, oracle):
    rng = np.random.default_rng(1)
    E = np.vstack([rng.normal(0, 0.2, (30, 9)), [a["e"] for a in oracle.TERMINAL.anchors[1:]]])
    B = E.shape[0]
    V = torch.empty(B, dtype=F64, device="cuda"); g = torch.empty(B, 9, dtype=F64, device="cuda")
    Hs = torch.empty(B, 81, dtype=F64, device="cuda")
    L.check(L.lib().ftmpc_terminal(eng6.handle, B, ptr(dev(E)), ptr(V), ptr(g), ptr(Hs), None))
    V, g, Hs = V.cpu().numpy(), g.cpu().numpy(), Hs.cpu().numpy()
    assert V[-2] == pytest.approx(82.584936488841, rel=1e-12) and V[-1] == pytest.approx(40.774178314212, rel=1e-12)
    for b in range(0, B, 5):
        assert V[b] == pytest.approx(float(oracle.TERMINAL.cost(E[b])), rel=1e-12, abs=1e-10)
        assert np.allclose(g[b], oracle.TERMINAL.grad(E[b]), rtol=1e-10, atol=1e-9)
        for i in range(9):
            d = np.zeros(9); d[i] = 1e-6
            fd = (oracle.TERMINAL.grad(E[b] + d) - oracle.TERMINAL.grad(E[b] - d)) / 2e-6
            assert np.allclose(Hs[b].reshape(9, 9)[i], fd, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("N", [8, 15, 20])      # 8: generic path, 15 / 20: register-tiled path
def test_condense_gauss_newton_vs_oracle(L, oracle, N):
    """K2: condensed Hessian / gradient at theta = 0 == 2R + sum_t G_t' 2Q G_t + G_N' Hq G_N from the
    oracle's complex-step sensitivities (Hq = quadratic part of V_f)"""
    eng = make_engine([[(10, 1.0), (11, 1.0)]], N)
    prob, _ = oracle.default_problem(N)
    rng = np.random.default_rng(7)
    U = rng.normal(0, 0.5, (N, 6))
    X, _ = prob.rollout(U)
    G, Hx = prob.sensitivities(X, U)
    W = U + np.concatenate([oracle.F_VIRT, np.zeros(3)])           # total wrench (hover: u_ref = 0)
    xd = dev(X[None].copy()); Wd = dev(W[None])
    jac = torch.zeros(1, N, 13, 13, dtype=F64, device="cuda")
    L.check(L.lib().ftmpc_rk4_jac(eng.handle, 1, ptr(xd), ptr(Wd), ptr(jac), None, None, None))
    eN = X[N, :9] - prob.xref[N]
    gV = oracle.TERMINAL.grad(eN)
    quad = np.array(eng.cfg.term_quad[:]).reshape(9, 9)
    n = 6 * N
    Hd = torch.zeros(1, n, n, dtype=F64, device="cuda"); gd = torch.zeros(1, n, dtype=F64, device="cuda")
    L.check(L.lib().ftmpc_condense(eng.handle, 1, ptr(jac), None, ptr(xd), ptr(dev(U[None])), ptr(dev(prob.xref[None])),
                                   ptr(dev(gV[None])), ptr(dev(quad[None])), C.c_double(0.0), ptr(Hd), ptr(gd), None))
    Q2, R2 = 2 * np.diag(prob.Q), 2 * np.diag(prob.R)
    Href = np.kron(np.eye(N), R2)
    gref = (2 * U * prob.R).ravel()
    for t in range(N):
        Href += G[t][:9].T @ Q2 @ G[t][:9]
        gref += G[t][:9].T @ (2 * prob.Q * (X[t, :9] - prob.xref[t]))
    Href += G[N][:9].T @ quad @ G[N][:9]
    gref += G[N][:9].T @ gV
    assert np.allclose(Hd.cpu().numpy()[0], Href, rtol=1e-10, atol=1e-9)
    assert np.allclose(gd.cpu().numpy()[0], gref, rtol=1e-10, atol=1e-9)
    f, grad, *_ = prob.fun_and_grad(U)
    assert np.allclose(gd.cpu().numpy()[0], grad, rtol=1e-9, atol=1e-8)       # == gradient of the NLP objective


def test_qp_solve_kkt(L, eng6):
    """K3: batched dense QP, sparse rows; solver-independent KKT check per instance"""
    rng = np.random.default_rng(4)
    n, m, B = 60, 90, 19
    Hs, gs, bs = [], [], []
    ptr_, idx, val = [0], [], []
    for i in range(m):
        cols = np.sort(rng.choice(n, 5, replace=False))
        idx += list(cols); val += list(rng.normal(size=5)); ptr_.append(len(idx))
    for _ in range(B):
        M = rng.normal(size=(n, n)); Hs.append(M @ M.T + n * np.eye(n)); gs.append(rng.normal(0, 8, n))
        bs.append(rng.uniform(0.0, 0.3, m))
    Hm, g, b = np.stack(Hs), np.stack(gs), np.stack(bs)
    x = torch.zeros(B, n, dtype=F64, device="cuda"); lam = torch.zeros(B, m, dtype=F64, device="cuda")
    st = torch.full((B,), -7, dtype=torch.int32, device="cuda")
    L.check(L.lib().ftmpc_qp_solve(eng6.handle, B, n, m, ptr(dev(Hm)), ptr(dev(g)), ptr(dev(ptr_, torch.int32)),
                                   ptr(dev(idx, torch.int32)), ptr(dev(val)), ptr(dev(b)), ptr(x), ptr(lam), ptr(st), None))
    x, lam, st = x.cpu().numpy(), lam.cpu().numpy(), st.cpu().numpy()
    assert (st == 0).all()
    Cm = np.zeros((m, n))
    for i in range(m):
        Cm[i, idx[ptr_[i]:ptr_[i + 1]]] = val[ptr_[i]:ptr_[i + 1]]
    for k in range(B):
        s = b[k] - Cm @ x[k]                                        # C x <= b
        assert s.min() > -1e-9 and lam[k].min() >= 0
        assert np.abs(Hm[k] @ x[k] + g[k] + Cm.T @ lam[k]).max() < 1e-8
        assert np.abs(lam[k] * s).max() < 1e-8
        assert (lam[k] > 0).sum() > 0                               # the problems are genuinely constrained


def test_allocate_vs_oracle(L, eng6, oracle):
    rng = np.random.default_rng(2)
    B = 70
    ub = np.full((B, 16), 3.4)
    for k in range(B):
        ub[k, rng.choice(16, k % 3, replace=False)] = 0.0
    udes = np.stack([oracle.D_ALLOC @ (rng.uniform(0, 3.4, 16) * (ub[k] > 0)) for k in range(B)])
    udes[-1] = [100.0, 0, 0, 0, 0, 0]                                # infeasible request
    th = torch.zeros(B, 16, dtype=F64, device="cuda"); st = torch.zeros(B, dtype=torch.int32, device="cuda")
    L.check(L.lib().ftmpc_allocate(eng6.handle, B, ptr(dev(udes)), ptr(dev(ub)), ptr(th), ptr(st), None))
    th, st = th.cpu().numpy(), st.cpu().numpy()
    assert (st[:-1] == 0).all() and st[-1] != 0
    assert np.allclose(th[:-1] @ oracle.D_ALLOC.T, udes[:-1], atol=1e-9)
    assert (th[:-1] >= 0).all() and (th[:-1] <= ub[:-1]).all()
    for k in range(0, B - 1, 7):
        tho, ok = oracle.allocate(udes[k], ub[k])
        assert ok and np.allclose(th[k], tho, atol=1e-7)


def test_plant_step_vs_oracle(L, oracle):
    eng = make_engine([[(10, 1.0), (11, 1.0)], [(2, 0.0)]], 5)
    from ft_mpc_b200.util import scenarios
    rng = np.random.default_rng(3)
    B = 65
    st = scenarios.random_states(B, 4)
    u = rng.uniform(0, 3.4, (B, 16))
    scen = np.arange(B) % 2
    noise = rng.uniform(0, 1e-3, (B, 13))
    sc = eng.scenario_tensors(dev(scen, torch.int64))
    raw = eng.plant_step(dev(st), dev(u), sc, None, normalize=False).cpu().numpy()
    full = eng.plant_step(dev(st), dev(u), sc, dev(noise), normalize=True).cpu().numpy()
    fsets = [oracle.FaultSet([(10, 1.0), (11, 1.0)]), oracle.FaultSet([(2, 0.0)])]
    for k in range(B):
        xo = oracle.plant_rk4(st[k], u[k], fsets[scen[k]], 0.1)
        assert np.allclose(raw[k], xo, rtol=1e-13, atol=1e-13)                               # == model.dynamics(x,u), sim_env.py:85
        assert np.allclose(full[k], oracle.normalize_quaternion_robot(xo + noise[k]), rtol=1e-13, atol=1e-13)   # :88-93


# ---------------------------------------------------------------------------------------------------------
# the whole step
# ---------------------------------------------------------------------------------------------------------
def run_cases(golden, ks, N, warm=False, **opts):
    sets, scen = H.gather_cases(golden, ks)
    eng = make_engine(sets, N, **opts)
    z = None
    out = eng.buffers(len(ks))
    if warm:
        out["z"].copy_(dev(np.stack([H.z_from_U0(golden["U0"][k, :N], N) for k in ks])))
    res = eng.step(dev(golden["x0"][ks]), dev(golden["xref"][ks][:, :N + 1]), dev(golden["uref"][ks][:, :N + 1]),
                   dev(scen, torch.int64), warm=warm, out=out)
    torch.cuda.synchronize()
    return eng, {k: v.cpu().numpy() for k, v in res.items() if k != "ws"}


@pytest.mark.parametrize("N", [15, 20])
def test_step_vs_golden(L, oracle, golden, N):
    """ftmpc_step on the golden instances: same KKT point as the oracle (u0 1e-5 rel, active set bit-exact)"""
    ks = [k for k in H.cases_with_horizon(golden, N) if not golden["warm"][k]]
    eng, out = run_cases(golden, ks, N)
    assert (out["status"] == 0).all(), out["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        name = golden["name"][k]
        u0 = golden["U"][k, 0]
        assert out["cost"][j] == pytest.approx(golden["f"][k], rel=1e-9), name
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), name
        assert np.array_equal(out["z"][j, :6], out["u0"][j])
        assert H.active_bits(out["active"][j].view(np.uint32), nbits) == H.active_bits(golden["active"][k], nbits), name
        assert np.allclose(out["thrust"][j], golden["thrust"][k], atol=2e-5), name
        # the returned decision vector is a rollout of the reference's dynamics: z = [u | x]
        prob = H.case_problem(oracle, golden, k)
        X, _ = prob.rollout(out["z"][j, :6 * N])
        assert np.allclose(out["z"][j, 6 * N:].reshape(N + 1, 13), X, rtol=1e-12, atol=1e-12), name


@pytest.mark.gpu
@pytest.mark.parametrize("N", [15, 20])
def test_step_accelerating_reference_vs_golden(L, oracle, accel_golden, N):
    """row f-3: circle references (non-zero nominal wrench, rotated per stage inside the kernels) against the oracle"""
    g = accel_golden
    ks = H.cases_with_horizon(g, N)
    assert len(ks) == 2 and np.abs(g["uref"][ks]).max() > 1.0
    eng, out = run_cases(g, ks, N)
    assert (out["status"] == 0).all(), out["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        name, u0 = g["name"][k], g["U"][k, 0]
        assert out["cost"][j] == pytest.approx(g["f"][k], rel=1e-9), name
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), name
        assert H.active_bits(out["active"][j].view(np.uint32), nbits) == H.active_bits(g["active"][k], nbits), name
        assert np.allclose(out["thrust"][j], g["thrust"][k], atol=2e-5), name
        kk = oracle.kkt_residual(H.case_problem(oracle, g, k), out["z"][j, :6 * N])
        assert kk["stat"] < 1e-5 and kk["viol"] < 1e-7, (name, kk["stat"], kk["viol"])


@pytest.mark.gpu
def test_step_on_bench_workload_vs_golden(L, oracle, bench_golden):
    """32 instances of bench.py's own workload against the oracle: u0 1e-5 relative, active sets bit-exact, thrust 2e-5"""
    g = bench_golden
    N = 20
    ks = H.cases_with_horizon(g, N)
    assert len(ks) == 32
    eng, out = run_cases(g, ks, N)
    assert (out["status"] == 0).all(), out["status"]
    nbits = 26 * N + 72
    for j, k in enumerate(ks):
        name, u0 = g["name"][k], g["U"][k, 0]
        assert out["cost"][j] == pytest.approx(g["f"][k], rel=1e-9), name
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max()), name
        assert H.active_bits(out["active"][j].view(np.uint32), nbits) == H.active_bits(g["active"][k], nbits), name
        assert np.allclose(out["thrust"][j], g["thrust"][k], atol=2e-5), name


@pytest.mark.gpu
def test_get_control_tracks_circle_reference(L, oracle, accel_golden):
    """the reference-facing call with an accelerating trajectory: load_trajectory("generate_circle") + get_control(x, t)
    reproduces the oracle's thrust for the golden window (circle0_N15 starts at step 5, t = 0.5), and the debug
    history records the step in the reference's export format"""
    from ft_mpc_b200.controllers import SpiralingController
    from ft_mpc_b200.models import SpiralModel, SystemModel
    from ft_mpc_b200.util import BrokenThruster, ControllerDebug
    g = accel_golden
    k = list(g["name"]).index("circle0_N15")
    model = SystemModel(0.1)
    for i, a in H.case_faults(g, k):
        model.set_fault(BrokenThruster(i, a))
    dbg = ControllerDebug()
    ctrl = SpiralingController(SpiralModel.from_system_model(model), {"horizon": 15}, dbg)
    ctrl.load_trajectory("generate_circle", 3)
    xr, ur = ctrl.get_next_trajectory_part(0.5)
    assert np.allclose(xr.T, g["xref"][k, :16], atol=1e-14) and np.allclose(ur.T, g["uref"][k, :16], rtol=1e-12, atol=1e-11)
    thrust = ctrl.get_control(g["x0"][k], 0.5)
    assert ctrl.last_status == 0
    assert np.abs(ctrl.last_u0 - g["U"][k, 0]).max() <= 1e-5 * max(1.0, np.abs(g["U"][k, 0]).max())
    assert np.allclose(thrust, g["thrust"][k], atol=2e-5)
    row = dbg.table()
    assert row.shape == (1, 67) and row[0, 0] == 0.5 and np.allclose(row[0, 14:30], thrust)
    assert np.allclose(row[0, 36:45], oracle.robot_to_center(g["x0"][k])[:9], atol=1e-12)          # circle state = c0[:9]
    assert np.allclose(row[0, 45:48], g["xref"][k, 0, 0:3] - g["x0"][k, 0:3])                       # position error


def test_step_kkt_residual_through_oracle(L, oracle, golden):
    """solver-independent check: the GPU's U* satisfies the KKT conditions of the oracle's restatement of the NLP"""
    N = 20
    ks = [k for k in H.cases_with_horizon(golden, N) if not golden["warm"][k]][:5]
    eng, out = run_cases(golden, ks, N)
    for j, k in enumerate(ks):
        r = oracle.kkt_residual(H.case_problem(oracle, golden, k), out["z"][j, :6 * N])
        # the SQP stops on the step (|d|_inf <= 1e-8, or <= 1e-6 once the predicted decrease is rounding noise); the
        # gradient residual is that times the Hessian scale (~1e2).  u0 itself is checked to 1e-5 in test_step_vs_golden.
        assert r["stat"] < 1e-4 and r["viol"] < 1e-8, (golden["name"][k], r["stat"], r["viol"])


def test_warm_start_closed_loop_vs_golden(L, oracle, golden):
    N = 15
    ks = [k for k in range(len(golden["N"])) if golden["warm"][k]]
    eng, out = run_cases(golden, ks, N, warm=True)
    assert (out["status"] == 0).all()
    for j, k in enumerate(ks):
        u0 = golden["U"][k, 0]
        assert np.abs(out["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max())
        assert np.allclose(out["thrust"][j], golden["thrust"][k], atol=2e-5)
        assert H.active_bits(out["active"][j].view(np.uint32), 462) == H.active_bits(golden["active"][k], 462)


def test_infeasible_instance_is_flagged(L, golden):
    k = list(golden["name"]).index("line_N15")
    eng, out = run_cases(golden, [k], 15)
    assert out["status"][0] != 0
    assert np.isfinite(out["thrust"]).all() and (out["thrust"] >= 0).all() and (out["thrust"] <= 3.4 + 1e-12).all()


def test_get_control_default_scenario_closed_loop(L, oracle, golden):
    """reference-facing API (sim.py:23-54, sim_env.py:77-99): SystemModel -> SpiralModel -> SpiralingController,
    get_control(x, t) in a closed loop with model.dynamics; parity with the golden closed-loop vectors"""
    from ft_mpc_b200.controllers import SpiralingController
    from ft_mpc_b200.models import SpiralModel, SystemModel
    from ft_mpc_b200.util import BrokenThruster
    model = SystemModel(0.1)
    model.set_fault(BrokenThruster(10, 1.0)); model.set_fault(BrokenThruster(11, 1.0))
    ctrl = SpiralingController(SpiralModel.from_system_model(model), {"horizon": 15, "param_set": "P1",
                               "P1": {"Q": [1, 1, 1, 1, 1, 1, 2, 2, 2], "R": [.1, .1, .1, .01, .01, .01]}}, None)
    ctrl.load_trajectory("hover", 30)
    names = list(golden["name"])
    x = golden["x0"][names.index("default_N15")].copy()
    t = 0.0
    for step in range(5):
        k = names.index("default_N15" if step == 0 else f"closed_loop_N15_step{step}")
        assert np.allclose(x, golden["x0"][k], atol=1e-6), step                 # closed-loop state within 1e-6
        u = ctrl.get_control(x, t)
        assert ctrl.last_status == 0
        assert u.shape == (16,) and u[10] == 0 and u[11] == 0 and (u >= 0).all() and (u <= 3.4).all()
        assert np.allclose(u, golden["thrust"][k], atol=2e-5), step
        x = model.normalize_quaternion(model.dynamics(x, u))
        t += 0.1


def test_gpu_matches_cpu_port(L, built, golden):
    """same algorithm, two instantiations (CudaBlock vs SerialBlock): results agree to rounding"""
    from ft_mpc_b200.util import scenarios
    N, B = 20, 48
    cells = scenarios.load_cells()
    eng = make_engine(cells, N)
    st = scenarios.random_states(B, 11)
    scen = (np.arange(B) * 7) % len(cells)
    xref = scenarios.hover_reference(B, N)
    out = eng.step(dev(st), dev(xref), scenario=dev(scen, torch.int64))
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in out.items() if k != "ws"}
    port = H.CpuPort()
    masks = np.array([eng.mask_tab[s] for s in scen], np.uint16)
    ffs = np.stack([eng.fault_force_tab[s] for s in scen])
    c = port.step(eng.cfg, eng.hull_table, st, xref, None, masks, ffs, scen)
    ok = (g["status"] == 0) & (c["status"] == 0)
    assert ok.sum() >= B - 1, (g["status"], c["status"])          # (all 48 converge on both; one straggler tolerated)
    # both stop within the SQP termination tolerance (step <= 1e-6 once the predicted decrease is rounding noise)
    assert np.abs(g["u0"] - c["u0"])[ok].max() < 2e-6 and np.abs(g["thrust"] - c["thrust"])[ok].max() < 2e-6
    assert np.array_equal(g["active"].view(np.uint32)[ok], c["active"][ok])


def test_qp_warm_start_gives_identical_results(L, golden):
    """warm_qp = 1 (working set of each QP started from the previous QP's active set) must not change the answer"""
    N = 20
    ks = [k for k in H.cases_with_horizon(golden, N) if not golden["warm"][k]]
    eng0, out0 = run_cases(golden, ks, N)
    eng1, out1 = run_cases(golden, ks, N, warm_qp=1)
    assert (out0["status"] == 0).all() and (out1["status"] == 0).all()
    assert np.abs(out0["u0"] - out1["u0"]).max() < 2e-6 and np.abs(out0["thrust"] - out1["thrust"]).max() < 2e-6
    assert np.array_equal(out0["active"], out1["active"])
    assert out1["iters"][:, 1].sum() < out0["iters"][:, 1].sum()          # fewer active-set iterations
    for j, k in enumerate(ks):
        u0 = golden["U"][k, 0]
        assert np.abs(out1["u0"][j] - u0).max() <= 1e-5 * max(1.0, np.abs(u0).max())


@pytest.mark.parametrize("N,B", [(30, 6), (100, 3)])
def test_long_horizon_global_scratch_path(L, built, N, B):
    """N = 30: the QP scratch (275 KB) no longer fits in shared memory -> k_solve<true> keeps it in a per-CTA global
    slice and the generic condensing / factorisation routines run; results must agree with the CPU port.
    N = 100 is the horizon of BASELINE configs[4] (600 decision variables, 2.9 MB of scratch per CTA)."""
    from ft_mpc_b200.util import scenarios
    cells = scenarios.load_cells(kinds=("single",))[:3]
    eng = make_engine(cells, N)
    st = scenarios.random_states(B, 21)
    scen = np.arange(B) % len(cells)
    xref = scenarios.hover_reference(B, N)
    out = eng.step(dev(st), dev(xref), scenario=dev(scen, torch.int64))
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in out.items() if k != "ws"}
    port = H.CpuPort()
    masks = np.array([eng.mask_tab[s] for s in scen], np.uint16)
    ffs = np.stack([eng.fault_force_tab[s] for s in scen])
    c = port.step(eng.cfg, eng.hull_table, st, xref, None, masks, ffs, scen)
    ok = (g["status"] == 0) & (c["status"] == 0)
    assert ok.sum() >= B - 1, (g["status"], c["status"])
    assert np.abs(g["u0"] - c["u0"])[ok].max() < 2e-6 and np.abs(g["thrust"] - c["thrust"])[ok].max() < 2e-6
    assert np.array_equal(g["active"].view(np.uint32)[ok], c["active"][ok])


def test_batch_1024_properties_and_sharding(L, oracle):
    """BASELINE config 3 (1,024 instances, N=20, single faults, random states): size-independent properties,
    and shard-invariance -- solving the batch in 2 / 8 contiguous shards gives bit-identical per-instance results"""
    from ft_mpc_b200.distributed import shard_bounds
    from ft_mpc_b200.util import scenarios
    N, B = 20, 1024
    cells = scenarios.load_cells(kinds=("single",))
    eng = make_engine(cells, N)
    st = scenarios.random_states(B, 0)
    scen = np.arange(B) % len(cells)
    xref = scenarios.hover_reference(B, N)
    std, xrd, scd = dev(st), dev(xref), dev(scen, torch.int64)
    out = eng.step(std, xrd, scenario=scd)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy().copy() for k, v in out.items() if k != "ws"}
    ok = g["status"] == 0
    assert ok.mean() > 0.995, np.bincount(g["status"], minlength=6)     # (the host build of the same solver converges on all 1024)
    th = g["thrust"]
    assert np.isfinite(th).all() and (th >= 0).all() and (th <= 3.4 + 1e-12).all()
    for k in range(B):                                                   # failed thrusters are never commanded
        for i, _ in cells[scen[k]]["faults"]:
            assert th[k, i] == 0.0
    # allocation reproduces the wrench the MPC asked for: D thrust = u0 + [f_virt;0] - D f_fault  (hover: u_ref = 0)
    for k in np.where(ok)[0][::16]:
        fs = oracle.FaultSet(cells[scen[k]]["faults"])
        u_des = g["u0"][k] + np.concatenate([oracle.F_VIRT, np.zeros(3)]) - fs.generalized
        assert np.allclose(oracle.D_ALLOC @ th[k], u_des, atol=1e-8)
    # KKT residual of a sample through the oracle
    for k in np.where(ok)[0][:: B // 6][:6]:
        fs = oracle.FaultSet(cells[scen[k]]["faults"])
        prob = oracle.Problem(fs, N, oracle.robot_to_center(st[k]), xref[k], np.zeros((N + 1, 6)))
        r = oracle.kkt_residual(prob, g["z"][k, :6 * N])
        assert r["stat"] < 1e-4 and r["viol"] < 1e-7, (k, r["stat"], r["viol"])     # step-based stop, see above
    # determinism + shard invariance
    for G in (2, 8):
        for r in range(G):
            lo, hi = shard_bounds(B, r, G)
            o2 = eng.step(std[lo:hi].contiguous(), xrd[lo:hi].contiguous(), scenario=scd[lo:hi].contiguous())
            torch.cuda.synchronize()
            for key in ("thrust", "u0", "status", "active", "iters", "cost"):
                assert np.array_equal(o2[key].cpu().numpy(), g[key][lo:hi]), (G, r, key)


def test_closed_loop_driver_matches_host_loop(L, oracle, golden):
    """BatchedMPC.closed_loop (device-side solve -> plant loop) == the golden closed loop of the oracle"""
    names = list(golden["name"])
    N = 15
    eng = make_engine([[(10, 1.0), (11, 1.0)]], N)
    traj, _ = oracle.assign_trajectory(oracle.hover_trajectory(30, 0.1), N, 0.1)
    x0 = dev(golden["x0"][[names.index("default_N15")]])
    xf, cost, worst, done = eng.closed_loop(x0, dev(traj.T.copy()), steps=4)
    assert int(worst.max()) == 0 and int(done[0]) == 4
    assert np.allclose(xf.cpu().numpy()[0], golden["x0"][names.index("closed_loop_N15_step4")], atol=1e-6)
    fsum = sum(golden["f"][names.index(n)] for n in ["default_N15"] + [f"closed_loop_N15_step{i}" for i in (1, 2, 3)])
    assert float(cost[0]) == pytest.approx(fsum, rel=1e-8)


@pytest.mark.gpu
def test_closed_loop_on_circle_matches_get_control_loop(L, oracle, accel_golden):
    """accelerating reference in closed loop: the device-side driver (windows of trajectory AND nominal wrench sliced on
    the device, warm start from the second step) reproduces the reference-style host loop get_control -> plant step
    (sim_env.py:77-99, noise off), whose first step is pinned by the oracle golden circle0_N15"""
    from ft_mpc_b200.controllers import SpiralingController
    from ft_mpc_b200.models import SpiralModel, SystemModel
    from ft_mpc_b200.util import BrokenThruster
    g = accel_golden
    k = list(g["name"]).index("circle0_N15")
    faults = H.case_faults(g, k)
    model = SystemModel(0.1)
    for i, a in faults:
        model.set_fault(BrokenThruster(i, a))
    ctrl = SpiralingController(SpiralModel.from_system_model(model), {"horizon": 15}, None)
    ctrl.load_trajectory("generate_circle", 3)
    fs = oracle.FaultSet(faults)
    x = g["x0"][k].copy()
    steps, k0 = 4, 5
    for j in range(steps):                                    # host loop, oracle plant
        thrust = ctrl.get_control(x, 0.1 * (k0 + j) + 1e-9)
        assert ctrl.last_status == 0
        if j == 0:
            assert np.allclose(thrust, g["thrust"][k], atol=2e-5)
        x = oracle.normalize_quaternion_robot(oracle.plant_rk4(x, thrust, fs, 0.1))
    eng = ctrl.engine
    xf, cost, worst, done = eng.closed_loop(dev(g["x0"][[k]]), ctrl._traj_dev, steps=steps, start_step=k0,
                                            nominal_input=ctrl._uref_dev)
    assert int(worst.max()) == 0 and int(done[0]) == steps
    assert np.allclose(xf.cpu().numpy()[0], x, atol=1e-6)
    # without the nominal wrench the loop ends somewhere else: the windows really are used
    xh, *_ = eng.closed_loop(dev(g["x0"][[k]]), ctrl._traj_dev, steps=steps, start_step=k0)
    assert np.abs(xh.cpu().numpy()[0] - x).max() > 1e-4


def test_hull_facets_kernel_vs_host_enumeration(L, eng6):
    """row f-2 on the device: ftmpc_hull_facets == the host facet enumeration (itself checked against the Qhull table of the
    reference's InputBounds in tests/test_abi_host.py) for tabulated cells, arbitrary intensities and the healthy vehicle;
    a rank-deficient pair is flagged; a controller built from the device table solves like one built from the host table"""
    from ft_mpc_b200 import _lib as LL
    from ft_mpc_b200.controllers.tools.input_bounds import zonotope_facets
    from ft_mpc_b200.models import SystemModel
    from ft_mpc_b200.util import scenarios
    m = SystemModel(0.1)
    cells = scenarios.load_cells(kinds=("single", "double"))
    rng = np.random.default_rng(5)
    sets = [c["faults"] for c in cells[::4]] + [[]] + [[(int(i), float(a))] for i, a in zip(rng.integers(0, 12, 6), rng.uniform(0, 1, 6))]
    sets += [[(2, 0.31), (9, 0.0)], [(12, 0.0), (13, 0.0)]]
    table, nrows, status = eng6.hull_facets(sets)
    torch.cuda.synchronize()
    table, nrows, status = table.cpu().numpy(), nrows.cpu().numpy(), status.cpu().numpy()
    assert status[-1] == 1                                          # pair (12,13): flat zonotope (Qhull raises for it)
    for k, fs in enumerate(sets[:-1]):
        A, b, r = zonotope_facets(m.D, m.max_thrust, fs)
        assert status[k] == 0 and r == 6 and nrows[k] == len(b), (fs, status[k], nrows[k], len(b))
        Ad = table[k, :LL.NH * LL.NU].reshape(LL.NH, LL.NU)
        bd = table[k, LL.NH * LL.NU:]
        assert np.allclose(Ad[:len(b)], A, atol=1e-12) and np.allclose(bd[:len(b)], b, atol=1e-11), fs
        assert np.all(Ad[len(b):] == 0) and np.all(bd[len(b):] == 1e30)
    # same KKT point from a controller that uses the device-built table (row order differs from Qhull's, the set does not)
    N = 15
    fs = [(10, 1.0), (11, 1.0)]
    ref = make_engine([fs], N)
    alt = make_engine([fs], N)
    t2, _, st2 = alt.hull_facets([fs])
    assert int(st2[0]) == 0
    import ctypes as C
    host_table = np.ascontiguousarray(t2.cpu().numpy())
    LL.lib().ftmpc_destroy(alt.handle)
    alt.handle = C.c_void_p()
    LL.check(LL.lib().ftmpc_create(C.byref(alt.handle), C.byref(alt.cfg), host_table.ctypes.data_as(C.POINTER(C.c_double))), "ftmpc_create")
    st = dev(scenarios.random_states(5, 3))
    xr = dev(scenarios.hover_reference(5, N))
    o1 = {k: v.clone() for k, v in ref.step(st, xr).items() if k != "ws"}
    o2 = alt.step(st, xr)
    torch.cuda.synchronize()
    ok = (o1["status"] == 0) & (o2["status"] == 0)
    assert int(ok.sum()) >= 4
    assert torch.allclose(o1["u0"][ok], o2["u0"][ok], atol=1e-7) and torch.allclose(o1["thrust"][ok], o2["thrust"][ok], atol=1e-6)


def test_concurrent_steps_on_two_streams(L):
    """steps issued with different buffer slots on different CUDA streams (bench.py pipelines consecutive batches that
    way: the work queue of a launch lives in its own workspace) give exactly the results of the same steps run one
    after the other"""
    from ft_mpc_b200.util import scenarios
    N, B = 20, 700                                   # > 148 CTAs: the dynamic queue is exercised
    cells = scenarios.load_cells(kinds=("single",))
    eng = make_engine(cells, N)
    xref = dev(scenarios.hover_reference(B, N))
    sa, sb = dev(scenarios.random_states(B, 31)), dev(scenarios.random_states(B, 32))
    sc = dev(np.arange(B) % len(cells), torch.int64)
    ref = []
    for st in (sa, sb):
        out = eng.step(st, xref, scenario=sc)
        torch.cuda.synchronize()
        ref.append({k: v.clone() for k, v in out.items() if k != "ws"})
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        o1 = eng.step(sa, xref, scenario=sc, out=eng.buffers(B, 1))
    with torch.cuda.stream(s2):
        o2 = eng.step(sb, xref, scenario=sc, out=eng.buffers(B, 2))
    torch.cuda.synchronize()
    for o, r in ((o1, ref[0]), (o2, ref[1])):
        for k in ("thrust", "u0", "status", "iters", "active", "cost"):
            assert torch.equal(o[k], r[k]), k
    assert int((ref[0]["status"] == 0).sum()) > 0.99 * B


def test_edge_cases_and_errors(L, oracle):
    from ft_mpc_b200.util import scenarios
    # horizon 1 and a single instance
    eng = make_engine([[(0, 0.0)]], 1)
    st = scenarios.random_states(1, 5)
    out = eng.step(dev(st), dev(scenarios.hover_reference(1, 1)))
    torch.cuda.synchronize()
    assert out["thrust"].shape == (1, 16) and torch.isfinite(out["thrust"]).all()
    # workspace too small / bad arguments -> error codes, no crash
    b = eng.buffers(1)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mask, ff, hidx = eng.scenario_tensors(torch.zeros(1, dtype=torch.int64, device="cuda"))
    args = [eng.handle, 1, ptr(dev(st)), ptr(dev(scenarios.hover_reference(1, 1))), None, ptr(mask), ptr(ff), ptr(hidx), 0,
            ptr(b["z"]), ptr(b["thrust"]), ptr(b["u0"]), ptr(b["active"]), ptr(b["status"]), ptr(b["iters"]), ptr(b["cost"])]
    assert L.lib().ftmpc_step(*args, ptr(b["ws"]), 16, stream) == -3
    assert L.lib().ftmpc_step(*args[:2], None, *args[3:], ptr(b["ws"]), b["ws"].numel(), stream) == -1
    assert L.lib().ftmpc_step(eng.handle, 0, *args[2:], ptr(b["ws"]), b["ws"].numel(), stream) == -1
    # an ill-posed fault cell (thruster 10 dead: f_virt outside the hull) is reported, never crashes
    eng = make_engine([[(10, 0.0)]], 10)
    st = scenarios.random_states(4, 6)
    out = eng.step(dev(st), dev(scenarios.hover_reference(4, 10)))
    torch.cuda.synchronize()
    th = out["thrust"].cpu().numpy()
    assert np.isfinite(th).all() and (th >= 0).all() and (th <= 3.4 + 1e-12).all() and (th[:, 10] == 0).all()
    assert set(out["status"].cpu().numpy().tolist()) <= {0, 1, 2, 3, 4}
