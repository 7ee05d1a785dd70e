"""Synthetic Monte-Carlo scenario batches: fault cells and initial robot states.

Host-side input generation only.  The hull of every fault cell (InputBounds, input_bounds.py:43-76) costs
~0.3 s of Qhull, so the table for all single and double faults x {dead, stuck-on} is cached in
data/hull_cells.npz (regenerate with `python -m ...scenarios` / tools/gen_hull_cells.py).
"""
from __future__ import annotations

import itertools
from pathlib import Path

import numpy as np

DATA = Path(__file__).resolve().parent.parent / "data" / "hull_cells.npz"
F_VIRT6 = np.array([0.0, 3.5, 0.0, 0.0, 0.0, 0.0])          # spiral_parameters.py:36


def all_cells():
    """16 singles + 120 pairs, each thruster dead (0.0) or stuck fully open (1.0)."""
    cells = []
    for i in range(16):
        for a in (0.0, 1.0):
            cells.append([(i, a)])
    for i, j in itertools.combinations(range(16), 2):
        for a, b in itertools.product((0.0, 1.0), repeat=2):
            cells.append([(i, a), (j, b)])
    return cells


def load_cells(min_margin: float = 1e-6, kinds=("single", "double")):
    """Well-posed cells: hull exists and f_virt is strictly inside it (margin = min(b - A f_virt) > min_margin).
    Returns list of dicts {faults, A, b, margin}."""
    z = np.load(DATA, allow_pickle=False)
    out = []
    for k in range(len(z["nfault"])):
        nf = int(z["nfault"][k])
        kind = "single" if nf == 1 else "double"
        if kind not in kinds or not z["ok"][k] or z["margin"][k] <= min_margin:
            continue
        faults = [(int(z["idx"][k, j]), float(z["inten"][k, j])) for j in range(nf)]
        nh = int(z["nh"][k])
        out.append(dict(faults=faults, A=z["A"][k, :nh].copy(), b=z["b"][k, :nh].copy(), margin=float(z["margin"][k])))
    return out


def random_states(B: int, seed: int) -> np.ndarray:
    """Robot states [p v q w]: pos U(-1,1)^3, vel U(-.5,.5)^3, euler zyx U(-pi,pi)^3, w = w_des + U(-.5,.5)^3."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1, 1, (B, 3))
    vel = rng.uniform(-0.5, 0.5, (B, 3))
    q = Rotation.from_euler("zyx", rng.uniform(-np.pi, np.pi, (B, 3))).as_quat()
    om = np.array([0.0, 0.0, 0.6]) + rng.uniform(-0.5, 0.5, (B, 3))
    return np.concatenate([pos, vel, q, om], axis=1)


def hover_reference(B: int, N: int, position=(0.0, 0.0, 0.0)) -> np.ndarray:
    """[B, N+1, 9] hover window: p_ref, v_ref = 0, w_ref = w_des (spiraling_mpc.py:269-277)."""
    w = np.zeros((B, N + 1, 9))
    w[..., 0:3] = position
    w[..., 8] = 0.6
    return w
