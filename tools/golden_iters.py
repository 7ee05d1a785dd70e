#!/usr/bin/env python
"""Development aid: iteration counts / final solver scalars of the GPU path on the golden instances.
python tools/golden_iters.py [max_sqp_iter] [case-name-for-scalar-dump]"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers as H
import ftmpc_import; ftmpc_import.load()
from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
from ft_mpc_b200.models import SystemModel
g = np.load(ROOT / "tests" / "golden" / "nlp_cases.npz")
mx = int(sys.argv[1]) if len(sys.argv) > 1 else 60
SC = ["F", "CSUM", "NU", "GD", "THETA", "DMAX", "DELTA", "LAMMAX", "STATUS", "ITER", "QPIT", "NACT", "CHOLFAIL", "QPST", "ALPHA", "CMAX", "SIGMA", "HFAIL"]
d = lambda a, dt=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
if len(sys.argv) > 2:
    k = list(g["name"]).index(sys.argv[2]); N = int(g["N"][k])
    n, nv, mc = 6 * N, 6 * N + 1, 26 * N + 72
    m = mc + 2
    osc = n + nv + 13 * (N + 1) + mc + 2 * m + 13 * (N + 1) + 2 * 169 * N + 9 + 81
    for it in list(range(1, mx + 1)):
        eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, [H.case_faults(g, k)], max_sqp_iter=it)
        out = eng.step(d(g["x0"][[k]]), d(g["xref"][[k]][:, :N + 1]), d(g["uref"][[k]][:, :N + 1]))
        torch.cuda.synchronize()
        sc = out["ws"].view(torch.float64)[osc:osc + len(SC)].cpu().numpy()
        print(f"it {it:2d} " + " ".join(f"{nm} {v:.3e}" for nm, v in zip(SC, sc) if nm in ("F", "CSUM", "THETA", "DMAX", "LAMMAX", "QPIT", "NACT", "CHOLFAIL", "ALPHA", "SIGMA", "GD", "STATUS")))
        if int(out["status"][0]) == 0:
            break
    sys.exit(0)
for N in (15, 20):
    ks = [k for k in H.cases_with_horizon(g, N) if not g["warm"][k]]
    sets, scen = H.gather_cases(g, ks)
    eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, sets, max_sqp_iter=mx)
    out = eng.step(d(g["x0"][ks]), d(g["xref"][ks][:, :N + 1]), d(g["uref"][ks][:, :N + 1]), d(scen, torch.int64))
    torch.cuda.synchronize()
    for j, k in enumerate(ks):
        print(N, g["name"][k], "status", int(out["status"][j]), "iters", out["iters"][j].tolist(), "cost", float(out["cost"][j]), "golden f", float(g["f"][k]))
