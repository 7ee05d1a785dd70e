#!/usr/bin/env python
"""Development aid: iteration counts of the GPU path on the golden instances (python tools/golden_iters.py [max_sqp_iter])."""
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
mx = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for N in (15, 20):
    ks = [k for k in H.cases_with_horizon(g, N) if not g["warm"][k]]
    sets, scen = H.gather_cases(g, ks)
    eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, sets, max_sqp_iter=mx)
    d = lambda a, dt=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
    out = eng.step(d(g["x0"][ks]), d(g["xref"][ks][:, :N + 1]), d(g["uref"][ks][:, :N + 1]), d(scen, torch.int64))
    torch.cuda.synchronize()
    for j, k in enumerate(ks):
        print(N, g["name"][k], "status", int(out["status"][j]), "iters", out["iters"][j].tolist(), "cost", float(out["cost"][j]), "golden f", float(g["f"][k]))
