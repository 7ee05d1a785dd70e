#!/usr/bin/env python
"""Tiny invocation of every hot-path kernel for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
4 instances at N = 20 (shared-memory path, hover + accelerating reference), 2 at N = 30 (global-scratch path), plant step,
device hull construction."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ftmpc_import; ftmpc_import.load()
from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
from ft_mpc_b200.models import SystemModel
from ft_mpc_b200.util import scenarios
d = lambda a, dt=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
cells = scenarios.load_cells(kinds=("single",))[:2]
for N, B in ((20, 4), (30, 2)):
    eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, cells)
    st = d(scenarios.random_states(B, 5)); xr = d(scenarios.hover_reference(B, N)); sc = d(np.arange(B) % 2, torch.int64)
    out = eng.step(st, xr, scenario=sc)
    torch.cuda.synchronize()
    print("N", N, "status", out["status"].tolist(), "iters", out["iters"].tolist())
    if N == 20:
        ur = torch.zeros(B, N + 1, 6, dtype=torch.float64, device="cuda"); ur[:, :, 0] = 1.0; ur[:, :, 1] = -0.5
        out = eng.step(st, xr, ur, scenario=sc)
        torch.cuda.synchronize()
        print("accelerating reference: status", out["status"].tolist())
        nxt = eng.plant_step(st, out["thrust"], sc)
        tab, nr, stt = eng.hull_facets([c["faults"] for c in cells])
        torch.cuda.synchronize()
        print("plant ok", bool(torch.isfinite(nxt).all()), "hull rows", nr.tolist(), stt.tolist())
