"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/ftmpc.h
declares, the ctypes struct mirrors the C struct, the product refuses to run without a GPU (no CPU
fallback), and the host-side mirror of the reference interface behaves like the reference's classes."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ftmpc.h"


def declared_symbols():
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ftmpc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(str(built[0]))
    syms = declared_symbols()
    assert len(syms) >= 15, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ftmpc.h but not exported by libftmpc.so"


def test_library_has_sm100a_code(built):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", str(built[0])], capture_output=True, text=True)
    assert "sm_100a" in out.stdout, out.stdout + out.stderr


def test_ctypes_struct_matches_c_struct(ft, tmp_path):
    """sizeof / field offsets of ftmpc_config as gcc sees them == the ctypes mirror in ft_mpc_b200._lib"""
    from ft_mpc_b200 import _lib as L
    fields = [f[0] for f in L.FtmpcConfig._fields_]
    src = tmp_path / "layout.c"
    body = "\n".join(f'printf("{f} %zu\\n", offsetof(ftmpc_config, {f}));' for f in fields)
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "ftmpc.h"\nint main(){{printf("sizeof %zu\\n", sizeof(ftmpc_config));\n{body}\nreturn 0;}}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    lines = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(lines["sizeof"]) == C.sizeof(L.FtmpcConfig)
    for f in fields:
        assert int(lines[f]) == getattr(L.FtmpcConfig, f).offset, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour without a GPU")
def test_no_cpu_fallback(ft, built):
    """without a CUDA device the library and the controller fail loudly"""
    import helpers as H
    from ft_mpc_b200 import _lib as L
    from ft_mpc_b200.controllers import SpiralingController
    from ft_mpc_b200.models import SpiralModel, SystemModel
    cfg, table, *_ = H.host_tables([[(10, 1.0), (11, 1.0)]], 15)
    h = C.c_void_p()
    rc = L.lib().ftmpc_create(C.byref(h), C.byref(cfg), table.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -5 and b"no CPU fallback" in L.lib().ftmpc_strerror(rc)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SpiralingController(SpiralModel.from_system_model(SystemModel(0.1)), {"horizon": 15})
    # the product package never imports the oracle
    for f in (ROOT / "fault-tolerant-mpc_b200").rglob("*.py"):
        assert "ftmpc_oracle" not in f.read_text() and "libftmpc_cpu" not in f.read_text(), f


def test_error_codes(ft, built):
    from ft_mpc_b200 import _lib as L
    lib = L.lib()
    assert lib.ftmpc_strerror(0) == b"ok"
    assert lib.ftmpc_create(None, None, None) == -1
    assert lib.ftmpc_num_var(None) == -1 and lib.ftmpc_workspace_bytes(None, 1, None) == -1
    with pytest.raises(RuntimeError, match="invalid argument"):
        L.check(-1, "x")


def test_model_mirror(ft, oracle):
    """SystemModel / SpiralModel / SpiralParameters carry the reference's constants (sys_model.py:52-131,228-243)"""
    from ft_mpc_b200.models import SpiralModel, SystemModel
    from ft_mpc_b200.util import BrokenThruster
    m = SystemModel(0.1)
    assert (m.Nx, m.Nu, m.Nu_simplified, m.Nu_full) == (13, 16, 6, 16) and m.mass == 16.8 and m.max_thrust == 3.4
    assert np.array_equal(m.D, oracle.D_ALLOC) and np.array_equal(np.diag(m.inertia), [0.2, 0.3, 0.25])
    m.set_fault(BrokenThruster(10, 1.0)); m.set_fault(BrokenThruster(11, 1.0))
    assert m.fault_mask == (1 << 10) | (1 << 11)
    assert np.allclose(m.faulty_force_generalized, [0, 6.8, 0, 0, 0, 0]) and m.u_ub_physical[10] == 0 and m.u_ub_physical[0] == 3.4
    s = SpiralModel.from_system_model(m)
    assert s.Nu == 6 and np.allclose(s.r, oracle.R_VEC) and len(s.broken_thrusters) == 2
    assert np.allclose(s.spiral_params.compensation_force, [0, -3.3, 0, 0, 0, 0])    # spiral_parameters.py:37 on the faulted model
    x = np.array([1, 0, 1, 1, .5, 0, 0.1, -0.2, 0.3, 0.9, 0.3, 0.8, -0.1])
    assert np.allclose(s.robot_to_center(x), oracle.robot_to_center(x), atol=1e-15)
    assert np.linalg.norm(m.normalize_quaternion(x)[6:10]) == pytest.approx(1.0)


def test_hull_table_and_scenarios(ft, oracle):
    from ft_mpc_b200._lib import HULL_STRIDE
    from ft_mpc_b200.controllers.spiraling_mpc import hull_table_entry
    from ft_mpc_b200.controllers.tools.input_bounds import hull_of_faults
    from ft_mpc_b200.util import scenarios
    A, b = hull_of_faults(oracle.D_ALLOC, 3.4, [(10, 1.0), (11, 1.0)])
    Ao, bo = oracle.input_bounds(oracle.FaultSet([(10, 1.0), (11, 1.0)]))
    assert np.array_equal(A, Ao) and np.array_equal(b, bo)          # same Qhull + np.unique row order (input_bounds.py:67-73)
    e = hull_table_entry(A[:18], b[:18])
    assert e.shape == (HULL_STRIDE,) and (e[26 * 6 + 18:] == 1e30).all() and (e[18 * 6:26 * 6] == 0).all()
    with pytest.raises(ValueError):
        hull_table_entry(np.zeros((27, 6)), np.zeros(27))
    cells = scenarios.load_cells()
    singles = scenarios.load_cells(kinds=("single",))
    assert len(singles) == 20 and len(cells) == 20 + 158                  # strictly feasible cells
    for c in singles:
        assert c["margin"] > 1e-6 and c["A"].shape[0] in (18, 26)
    # cached table == fresh Qhull for one cell
    c = cells[33]
    A2, b2 = hull_of_faults(oracle.D_ALLOC, 3.4, c["faults"])
    assert np.allclose(A2, c["A"]) and np.allclose(b2, c["b"])
    st = scenarios.random_states(16, 0)
    assert st.shape == (16, 13) and np.allclose(np.linalg.norm(st[:, 6:10], axis=1), 1.0)
    assert np.array_equal(st, scenarios.random_states(16, 0))


def test_trajectory_window_matches_oracle(ft, oracle):
    """load_trajectory + assign_trajectory + window slicing (spiraling_mpc.py:240-286,356-365), host side only"""
    from ft_mpc_b200.util.get_trajectory import load_trajectory
    tr = load_trajectory("hover", 0.1, 30)
    assert tr.shape == (13, 3000) and np.array_equal(tr, oracle.hover_trajectory(30, 0.1))
    traj, nom = oracle.assign_trajectory(tr, 15, 0.1)
    assert traj.shape == (9, 3015) and np.allclose(traj[6:9].T, [0, 0, 0.6]) and np.abs(nom).max() == 0
    c = load_trajectory("circle_r_2_sPerFullCircle_30", 0.1, 3)
    assert np.allclose(c[0:3, 0], 0) and np.allclose(np.hypot(c[0] + 2, c[1]), 2)
    with pytest.raises(ValueError):
        load_trajectory("nonsense", 0.1, 1)


def test_shard_bounds(ft):
    from ft_mpc_b200.distributed import shard_bounds
    for B in (1, 7, 1024, 65536, 65537):
        for G in (1, 2, 4, 8):
            blocks = [shard_bounds(B, r, G) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _gloo_worker(rank, world, port, batch, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200.distributed import gather_results, pack_results, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(batch, rank, world)
    ids = torch.arange(lo, hi, dtype=torch.float64)
    rec = pack_results(ids[:, None] * torch.ones(1, 13, dtype=torch.float64), ids * 2, (ids % 5).to(torch.int32),
                       torch.full((hi - lo,), 300, dtype=torch.int32))
    full = gather_results(rec, batch)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [10, 11])
def test_all_gather_results_world2_gloo(batch):
    """N>1 path on CPU: two gloo ranks shard a batch, all-gather the result records, order is preserved"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ids = np.arange(batch, dtype=float)
    assert full.shape == (batch, 16)
    assert np.array_equal(full[:, 0], ids) and np.array_equal(full[:, 13], 2 * ids)
    assert np.array_equal(full[:, 14], ids % 5) and (full[:, 15] == 300).all()


def test_analytic_zonotope_facets_match_qhull_table(ft):
    """Row f-2: facets of the wrench zonotope by direct enumeration == the Qhull hulls of the reference's InputBounds
    (input_bounds.py:43-76) for every tabulated single / double fault cell, as SETS of rows (the reference's row order
    is an artefact of Qhull's rounding noise); rank-deficient cells, where Qhull raises, come back as flat polytopes."""
    from ft_mpc_b200.controllers.tools.input_bounds import hull_of_faults, zonotope_facets
    from ft_mpc_b200.models import SystemModel
    from ft_mpc_b200 import _lib as L
    z = np.load(L.DATA_DIR / "hull_cells.npz")
    m = SystemModel(0.1)
    checked = 0
    for c in range(0, len(z["nh"]), 3):                       # every third cell keeps the CPU suite short
        nh = int(z["nh"][c])
        if nh == 0:
            continue
        faults = [(int(z["idx"][c, j]), float(z["inten"][c, j])) for j in range(int(z["nfault"][c]))]
        A, b, r = zonotope_facets(m.D, m.max_thrust, faults)
        assert r == 6 and A.shape == (nh, 6)
        assert np.allclose(np.linalg.norm(A, axis=1), 1.0, atol=1e-12)
        got = {tuple(x) for x in np.round(np.column_stack([A, b]), 7) + 0.0}
        want = {tuple(x) for x in np.round(np.column_stack([z["A"][c, :nh], z["b"][c, :nh]]), 7) + 0.0}
        assert got == want, faults
        checked += 1
    assert checked > 100
    # canonical order is deterministic and independent of the intensities (A depends on the mask only)
    A0, b0 = hull_of_faults(m.D, m.max_thrust, [(3, 0.0)], method="analytic")
    A1, b1 = hull_of_faults(m.D, m.max_thrust, [(3, 0.37)], method="analytic")
    assert np.array_equal(A0, A1) and np.allclose(b1 - b0, A0 @ (m.D[:, 3] * 0.37 * m.max_thrust), atol=1e-12)
    # every corner wrench satisfies the rows, and each row is tight for some corner
    rng = np.random.default_rng(0)
    u = rng.integers(0, 2, (400, 16)) * m.max_thrust
    u[:, 3] = 0.37 * m.max_thrust
    s = (m.D @ u.T).T @ A1.T - b1
    assert s.max() <= 1e-9
    # rank-deficient pair (Qhull raises QhullError for it): flat polytope, opposite rows with zero width
    A, b, r = zonotope_facets(m.D, m.max_thrust, [(12, 0.0), (13, 0.0)])
    assert r == 5 and np.isfinite(A).all()
    flat = [(i, j) for i in range(len(A)) for j in range(i) if np.allclose(A[i], -A[j], atol=1e-9) and abs(b[i] + b[j]) < 1e-9]
    assert len(flat) == 1


def test_bench_reference_arm_prints_one_json_line(built):
    """`bench.py --impl reference` (the driver's CPU arm): exactly one line on stdout, the contract's keys, no GPU needed"""
    import json
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "8"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0


def test_state_bounds_are_refused_with_the_reference_lines():
    """params['xub'] / params['xlb'] (spiraling_mpc.py:129-130, 180-185) are part of the reference constructor's contract and
    are REFUSED here -- before any device work, so the refusal does not depend on a GPU"""
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200.controllers import SpiralingController
    from ft_mpc_b200.models import SpiralModel, SystemModel
    model = SpiralModel.from_system_model(SystemModel(0.1))
    for key in ("xub", "xlb"):
        with pytest.raises(NotImplementedError, match="spiraling_mpc.py:129-130"):
            SpiralingController(model, {"horizon": 15, key: [1.0] * 13}, None)
