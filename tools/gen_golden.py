#!/usr/bin/env python
"""Generate tests/golden/nlp_cases.npz: seeded MPC instances solved by the CPU oracle (oracle/ftmpc_oracle.py,
scipy SLSQP + Newton polish on the reference NLP) to tight KKT tolerance.

The reference's own solver stack (casadi/IPOPT, cvxpy/OSQP) is not installable in this container, so these
vectors are oracle outputs, not reference outputs (see the oracle header: "parity unpinned" for the NLP solve).
Each case stores the inputs of one get_control call and the oracle's KKT point:
    faults (idx,intensity) x2 (idx -1 = unused), N, robot state x0[13], xref[N+1,9], uref[N+1,6], warm U0
    U*[N,6], f*, active rows (bit mask), kkt residuals, u_res, u_des, thrust[16], alloc_ok
Usage: python tools/gen_golden.py [nproc]         (about 10 minutes on 8 cores)
"""
import sys
import time
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import ftmpc_import  # noqa: E402

ftmpc_import.load()
import ftmpc_oracle as o  # noqa: E402
from ft_mpc_b200.util import scenarios  # noqa: E402

NMAX = 20


def make_cases():
    cases = []
    # (1) examples/sim.py default scenario, cold start, N = 15 (reference default) and N = 20
    for N in (15, 20):
        prob, x0 = o.default_problem(N)
        cases.append(dict(name=f"default_N{N}", faults=[(10, 1.0), (11, 1.0)], N=N, x0=x0, xref=prob.xref, uref=prob.uref))
    # (2) config-3 style: N = 20, well-posed single faults cycled, seeded random states, hover
    singles = scenarios.load_cells(kinds=("single",))
    st = scenarios.random_states(12, 0)
    for i in range(12):
        c = singles[i % len(singles)]
        cases.append(dict(name=f"single_N20_{i}", faults=c["faults"], N=20, x0=st[i],
                          xref=scenarios.hover_reference(1, 20)[0], uref=np.zeros((21, 6))))
    # (3) config-4 style: double faults, N = 15
    doubles = scenarios.load_cells(kinds=("double",))
    st = scenarios.random_states(6, 1)
    for i in range(6):
        c = doubles[(37 * i + 5) % len(doubles)]
        cases.append(dict(name=f"double_N15_{i}", faults=c["faults"], N=15, x0=st[i],
                          xref=scenarios.hover_reference(1, 15)[0], uref=np.zeros((16, 6))))
    # (4) constant-velocity line reference (get_trajectory.generate_line): zero nominal wrench, moving target
    for i, N in enumerate((15, 20)):
        line = np.zeros((13, 400)); line[0] = 0.1 * np.arange(400) * 0.1 * 10; line[3] = 1.0; line[9] = 1.0
        traj, nom = o.assign_trajectory(line, N, 0.1)
        xr, ur = o.window(traj, nom, 3, N)
        x0 = scenarios.random_states(2, 7)[i]
        cases.append(dict(name=f"line_N{N}", faults=[(3, 0.0)], N=N, x0=x0, xref=xr, uref=ur))
    return cases


def solve_case(c):
    t0 = time.time()
    fs = o.FaultSet(list(c["faults"]))
    prob = o.Problem(fs, c["N"], o.robot_to_center(c["x0"]), c["xref"], c["uref"])
    U0 = c.get("U0")
    res = o.get_control(prob, o.solve_nlp(prob, U0=U0))
    return c["name"], res, time.time() - t0


def pack(cases, results):
    K = len(cases)
    out = dict(name=np.array([c["name"] for c in cases]), N=np.zeros(K, np.int32), fault_idx=-np.ones((K, 2), np.int32),
               fault_inten=np.zeros((K, 2)), x0=np.zeros((K, 13)), xref=np.zeros((K, NMAX + 1, 9)),
               uref=np.zeros((K, NMAX + 1, 6)), warm=np.zeros(K, np.int32), U0=np.zeros((K, NMAX, 6)),
               U=np.zeros((K, NMAX, 6)), f=np.zeros(K), kkt_stat=np.zeros(K), kkt_viol=np.zeros(K),
               polished=np.zeros(K, bool), active=np.zeros((K, (26 * NMAX + 72 + 31) // 32), np.uint32),
               n_h=np.zeros(K, np.int32), u_res=np.zeros((K, 6)), u_des=np.zeros((K, 6)), thrust=np.zeros((K, 16)),
               alloc_ok=np.zeros(K, bool))
    for k, (c, r) in enumerate(zip(cases, results)):
        N = c["N"]
        out["N"][k] = N
        for j, (i, a) in enumerate(c["faults"]):
            out["fault_idx"][k, j], out["fault_inten"][k, j] = i, a
        out["x0"][k] = c["x0"]
        out["xref"][k, :N + 1] = c["xref"]
        out["uref"][k, :N + 1] = c["uref"]
        if c.get("U0") is not None:
            out["warm"][k] = 1
            out["U0"][k, :N] = np.asarray(c["U0"]).reshape(N, 6)
        out["U"][k, :N] = r["U"]
        out["f"][k], out["kkt_stat"][k], out["kkt_viol"][k], out["polished"][k] = r["f"], r["kkt_stat"], r["kkt_viol"], r["polished"]
        fsn = o.Problem(o.FaultSet(list(c["faults"])), N, o.robot_to_center(c["x0"]), c["xref"], c["uref"]).n_h
        out["n_h"][k] = fsn
        # active rows in the library's 26-rows-per-stage layout (cells with n_h < 26 are zero-padded per stage)
        for row in r["active"]:
            if row < fsn * N:
                t, i = divmod(int(row), fsn)
                bit = 26 * t + i
            else:
                bit = 26 * N + int(row) - fsn * N
            out["active"][k, bit // 32] |= np.uint32(1 << (bit % 32))
        out["u_res"][k], out["u_des"][k], out["thrust"][k], out["alloc_ok"][k] = r["u_res"], r["u_des"], r["thrust"], r["alloc_ok"]
    return out


def main():
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    cases = make_cases()
    with Pool(nproc) as p:
        results = {}
        for name, res, dt in p.imap_unordered(solve_case, cases):
            results[name] = res
            print(f"{name}: f={res['f']:.6f} kkt={res['kkt_stat']:.1e} viol={res['kkt_viol']:.1e} polished={res['polished']} "
                  f"nact={len(res['active'])} {dt:.0f}s", flush=True)
    # (5) closed loop, warm start: default scenario, 4 steps with the oracle plant (noise off), N = 15
    prob, x = o.default_problem(15)
    fs = prob.fs
    traj, nom = o.assign_trajectory(o.hover_trajectory(30, 0.1), 15, 0.1)
    Uprev = results["default_N15"]["U"]
    thrust = results["default_N15"]["thrust"]
    for step in range(1, 5):
        x = o.normalize_quaternion_robot(o.plant_rk4(x, thrust, fs, 0.1))
        xr, ur = o.window(traj, nom, step, 15)
        U0 = np.vstack([Uprev[1:], np.zeros((1, 6))])          # shifted previous solution (spiraling_mpc.py:324-331)
        c = dict(name=f"closed_loop_N15_step{step}", faults=[(10, 1.0), (11, 1.0)], N=15, x0=x.copy(), xref=xr, uref=ur, U0=U0)
        name, res, dt = solve_case(c)
        print(f"{name}: f={res['f']:.6f} kkt={res['kkt_stat']:.1e} {dt:.0f}s", flush=True)
        cases.append(c)
        results[name] = res
        Uprev, thrust = res["U"], res["thrust"]
    out = pack(cases, [results[c["name"]] for c in cases])
    dst = ROOT / "tests" / "golden" / "nlp_cases.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, "cases", len(cases))


def make_accel_cases():
    """(6) accelerating references: circles of get_trajectory.generate_circle, non-zero nominal
    wrench rotated into the body frame per stage (spiraling_mpc.py:156-166, 279-286)."""
    from ft_mpc_b200.util.get_trajectory import load_trajectory
    cases = []
    st = scenarios.random_states(4, 11)
    specs = [("generate_circle", 15, [(10, 1.0), (11, 1.0)], 5), ("generate_circle", 20, [(3, 0.0)], 12),
             ("circle_r_1_sPerFullCircle_15", 15, [(3, 0.0)], 7), ("circle_r_1_sPerFullCircle_15", 20, [(10, 1.0), (11, 1.0)], 30)]
    for i, (cmd, N, faults, k0) in enumerate(specs):
        traj, nom = o.assign_trajectory(load_trajectory(cmd, 0.1, 3), N, 0.1)
        xr, ur = o.window(traj, nom, k0, N)
        x0 = st[i].copy()
        x0[0:3] = 0.5 * x0[0:3] + xr[0, 0:3]                 # start in the neighbourhood of the moving target
        cases.append(dict(name=f"circle{i}_N{N}", faults=faults, N=N, x0=x0, xref=xr, uref=ur))
    return cases


def main_accel():
    cases = make_accel_cases()
    with Pool(4) as p:
        results = {}
        for name, res, dt in p.imap_unordered(solve_case, cases):
            results[name] = res
            print(f"{name}: f={res['f']:.6f} kkt={res['kkt_stat']:.1e} viol={res['kkt_viol']:.1e} polished={res['polished']} "
                  f"nact={len(res['active'])} {dt:.0f}s", flush=True)
    out = pack(cases, [results[c["name"]] for c in cases])
    dst = ROOT / "tests" / "golden" / "nlp_cases_accel.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, "cases", len(cases), "max |uref|", np.abs(out["uref"]).max())


def make_more_cases():
    """(7) a wider sample of the bench workload itself (BASELINE configs[3]): 32 instances of bench.make_workload at N = 20,
    i.e. single and double faults, dead and stuck-on thrusters, the bench's state distribution -- cold start, hover."""
    sys.path.insert(0, str(ROOT))
    import bench
    cells, states, scen, xref = bench.make_workload(8192, 20, 1)
    cases = []
    for k in list(range(0, 4096, 128)):
        cases.append(dict(name=f"bench_{k}", faults=cells[scen[k]]["faults"], N=20, x0=states[k], xref=xref[k], uref=np.zeros((21, 6))))
    return cases


def main_more():
    cases = make_more_cases()
    with Pool(7) as p:
        results = {}
        for name, res, dt in p.imap_unordered(solve_case, cases):
            results[name] = res
            print(f"{name}: f={res['f']:.6f} kkt={res['kkt_stat']:.1e} viol={res['kkt_viol']:.1e} polished={res['polished']} "
                  f"nact={len(res['active'])} {dt:.0f}s", flush=True)
    out = pack(cases, [results[c["name"]] for c in cases])
    dst = ROOT / "tests" / "golden" / "nlp_cases_bench.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, "cases", len(cases), "with KKT point", int((out["kkt_viol"] < 1e-8).sum()))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "--more":
    main_more()
elif __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "--accel":
    main_accel()
elif __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1].startswith("--")):
    main()


def refresh_active():
    """Recompute only the active-set masks / facet counts from the stored KKT points (used after the constraint ROW ORDER of
    the hull changed to the reference's: the minimisers themselves do not depend on the numbering)."""
    dst = ROOT / "tests" / "golden" / "nlp_cases.npz"
    g = dict(np.load(dst))
    K = len(g["N"])
    g["active"] = np.zeros_like(g["active"])
    for k in range(K):
        N = int(g["N"][k])
        faults = [(int(i), float(a)) for i, a in zip(g["fault_idx"][k], g["fault_inten"][k]) if i >= 0]
        prob = o.Problem(o.FaultSet(faults), N, o.robot_to_center(g["x0"][k]), g["xref"][k, :N + 1].copy(), g["uref"][k, :N + 1].copy())
        r = o.kkt_residual(prob, g["U"][k, :N].ravel())
        assert g["kkt_viol"][k] > 1e-8 or (r["stat"] < 1e-8 and r["viol"] < 1e-8), (g["name"][k], r["stat"], r["viol"])
        g["n_h"][k] = prob.n_h
        for row in r["active"]:
            if row < prob.n_h * N:
                t, i = divmod(int(row), prob.n_h)
                bit = 26 * t + i
            else:
                bit = 26 * N + int(row) - prob.n_h * N
            g["active"][k, bit // 32] |= np.uint32(1 << (bit % 32))
        print(g["name"][k], "active rows", len(r["active"]), flush=True)
    np.savez_compressed(dst, **g)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "--refresh-active":
    refresh_active()
