#!/usr/bin/env python
"""The bench scenarios that end with status = 1 (iteration cap / stall): does the ORACLE find a KKT point for them, and is it
the point the solver was heading for?   python tools/hard_instances.py [ids ...]   (CPU only; a few minutes)"""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tools"))
import ftmpc_oracle as o
import cpu_experiment as ce
import bench

ids = [int(a) for a in sys.argv[1:]] or [249, 2829, 6736]
N = 20
cells, states, scen, xref = bench.make_workload(8192, N, 1)
for i in ids:
    out = ce.run(1, N, {}, i)
    out_long = ce.run(1, N, {"max_sqp_iter": 400, "stall_window": 0}, i)
    fs = o.FaultSet(cells[scen[i]]["faults"])
    prob = o.Problem(fs, N, o.robot_to_center(states[i]), xref[i], np.zeros((N + 1, 6)))
    t0 = time.time()
    sol = o.solve_nlp(prob)
    k = o.kkt_residual(prob, sol["U"])
    Uc = out_long["z"][0, :6 * N]
    kc = o.kkt_residual(prob, Uc)
    print(f"instance {i}: faults {cells[scen[i]]['faults']}")
    print(f"  solver (default caps): status {int(out['status'][0])}, {int(out['iters'][0, 0])} SQP iterations, cost {out['cost'][0]:.9f}")
    print(f"  solver (400 iterations, no stall detector): status {int(out_long['status'][0])}, {int(out_long['iters'][0, 0])} iterations, "
          f"cost {out_long['cost'][0]:.9f}, oracle KKT residual of that point: stat {kc['stat']:.2e} viol {kc['viol']:.2e}")
    print(f"  oracle (SLSQP + polish, {time.time() - t0:.0f} s): f {sol['f']:.9f}, KKT stat {k['stat']:.2e} viol {k['viol']:.2e}, "
          f"|U_oracle - U_solver|_inf {np.abs(sol['U'].ravel() - Uc).max():.3e}")
