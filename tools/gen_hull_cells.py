#!/usr/bin/env python
"""Precompute the input-bound hull (InputBounds, input_bounds.py:43-76) of every single/double fault cell
-> fault-tolerant-mpc_b200/data/hull_cells.npz.  Uses the product's own construct-time routine."""
import sys
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ftmpc_import  # noqa: E402

ftmpc_import.load()
from ft_mpc_b200.controllers.tools.input_bounds import hull_of_faults  # noqa: E402
from ft_mpc_b200.models.sys_model import allocation_matrix  # noqa: E402
from ft_mpc_b200.util import scenarios  # noqa: E402


def one(fs):
    try:
        A, b = hull_of_faults(allocation_matrix(), 3.4, fs)
        return True, A, b
    except Exception:
        return False, None, None


if __name__ == "__main__":
    cells = scenarios.all_cells()
    with Pool(8) as p:
        res = p.map(one, cells, chunksize=4)
    K = len(cells)
    A = np.zeros((K, 26, 6)); b = np.zeros((K, 26)); nh = np.zeros(K, np.int32); ok = np.zeros(K, bool)
    margin = np.full(K, -np.inf); idx = np.zeros((K, 2), np.int32); inten = np.zeros((K, 2)); nfault = np.zeros(K, np.int32)
    for k, (fs, (good, Ak, bk)) in enumerate(zip(cells, res)):
        nfault[k] = len(fs)
        for j, (i, a) in enumerate(fs):
            idx[k, j], inten[k, j] = i, a
        if good and len(bk) <= 26:
            ok[k] = True; nh[k] = len(bk); A[k, :len(bk)] = Ak; b[k, :len(bk)] = bk
            margin[k] = np.min(bk - Ak @ scenarios.F_VIRT6)
    np.savez_compressed(scenarios.DATA, A=A, b=b, nh=nh, ok=ok, margin=margin, idx=idx, inten=inten, nfault=nfault)
    print("cells", K, "ok", ok.sum(), "strictly feasible", (margin > 1e-6).sum(), "facet counts", np.unique(nh[ok], return_counts=True))
