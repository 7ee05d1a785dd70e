#!/usr/bin/env python
"""Development aid: run a sample of the bench workload through the CPU checker (oracle/_cpu) with ftmpc_config
overrides and report iteration statistics.  Never part of the product path.
    python tools/cpu_experiment.py [--n 256] [--opts '{"sqp_tol": 1e-8}'] [--ref-opts '{...}']
With --ref-opts the same sample is solved a second time and u0 / active sets are compared."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers as H
import bench
import ftmpc_import; ftmpc_import.load()
from ft_mpc_b200 import _lib as L
from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, hull_table_entry
from ft_mpc_b200.controllers.tools.spiral_parameters import SpiralParameters
from ft_mpc_b200.models import SystemModel


def run(n, N, opts, offset=0):
    cells, states, scen, xref = bench.make_workload(8192, N, 1)
    port = H.CpuPort()
    model = SystemModel(0.1)
    sp = SpiralParameters(model)
    table = np.ascontiguousarray(np.stack([hull_table_entry(c["A"], c["b"]) for c in cells]))
    cfg = L.make_config(N, DEFAULT_Q, DEFAULT_R, dt=model.dt, mass=model.mass, inertia=model.inertia, r=sp.r, f_virt=sp.f_virt,
                        max_thrust=model.max_thrust, D=model.D, n_hull_sets=len(cells), **opts)
    masks = np.zeros(len(cells), np.uint16)
    ffs = np.zeros((len(cells), 16))
    for k, c in enumerate(cells):
        for i, a in c["faults"]:
            masks[k] |= np.uint16(1 << i)
            ffs[k, i] = a * model.max_thrust
    s = slice(offset, offset + n)
    t0 = time.perf_counter()
    out = port.step(cfg, table, states[s], xref[s], None, masks[scen[s]], ffs[scen[s]], scen[s], 0, None, 0)
    out["dt"] = time.perf_counter() - t0
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--N", type=int, default=20)
    ap.add_argument("--offset", type=int, default=0)
    ap.add_argument("--opts", default="{}")
    ap.add_argument("--ref-opts", default=None)
    a = ap.parse_args()
    out = run(a.n, a.N, json.loads(a.opts), a.offset)
    it = out["iters"]
    print(f"opts {a.opts}: ok {np.mean(out['status'] == 0):.4f} status hist {np.bincount(out['status'], minlength=5).tolist()} "
          f"sqp mean {it[:, 0].mean():.3f} max {it[:, 0].max()} qp mean {it[:, 1].mean():.2f}  ({out['dt']:.1f} s)")
    if a.ref_opts is not None:
        ref = run(a.n, a.N, json.loads(a.ref_opts), a.offset)
        ok = (out["status"] == 0) & (ref["status"] == 0)
        du = np.abs(out["u0"] - ref["u0"]).max(axis=1) / np.maximum(1.0, np.abs(ref["u0"]).max(axis=1))
        same = (out["active"] == ref["active"]).all(axis=1)
        rit = ref["iters"]
        print(f"ref  {a.ref_opts}: ok {np.mean(ref['status'] == 0):.4f} sqp mean {rit[:, 0].mean():.3f} qp mean {rit[:, 1].mean():.2f}")
        print(f"both ok {ok.sum()}: max rel du0 {du[ok].max():.3e}, active sets identical {same[ok].mean():.4f}, "
              f"thrust max diff {np.abs(out['thrust'] - ref['thrust'])[ok].max():.3e}")
