"""Pinning against the REFERENCE'S OWN CODE: tests/golden/ref_fixtures.npz holds outputs of ft_mpc's model modules
(ft_mpc.util.utils, ft_mpc.models.sys_model / spiral_model, SpiralParameters, InputBounds, get_trajectory) executed
in the build container by tools/gen_ref_fixtures.py.  The oracle, the host mirror and (with -m gpu) the CUDA kernels
must reproduce them.  What is not pinned this way is the NLP solve itself (IPOPT) and the allocator (CVXPY/OSQP)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
TAGS = {"default": [(10, 1.0), (11, 1.0)], "single3dead": [(3, 0.0)], "nofault": []}


@pytest.fixture(scope="module")
def ref():
    return np.load(ROOT / "tests" / "golden" / "ref_fixtures.npz")


def test_oracle_rotations_match_reference(oracle, ref):
    q = ref["quat"]
    assert np.allclose(oracle.rot(q), ref["Rot"], rtol=0, atol=1e-15)                     # utils.py:4-19
    assert np.allclose(np.swapaxes(oracle.rot(q), -1, -2), ref["RotInv"], rtol=0, atol=1e-15)   # utils.py:21-31
    full = np.zeros((len(q), 6, 6)); full[:, :3, :3] = oracle.rot(q); full[:, 3:, 3:] = np.eye(3)
    assert np.allclose(full, ref["RotFull"], atol=1e-15) and np.allclose(np.swapaxes(full, -1, -2), ref["RotFullInv"], atol=1e-15)


@pytest.mark.parametrize("tag", list(TAGS))
def test_oracle_model_matches_reference(oracle, ref, tag):
    fs = oracle.FaultSet(TAGS[tag])
    assert np.array_equal(oracle.D_ALLOC, ref[f"{tag}::D"])                                  # sys_model.py:73-123
    assert np.allclose([oracle.MASS, oracle.MAX_THRUST, *np.diag(oracle.INERTIA)], ref[f"{tag}::consts"], atol=0)
    assert np.array_equal(fs.faulty_force, ref[f"{tag}::faulty_force"])                      # sys_model.py:239
    assert np.allclose(fs.generalized, ref[f"{tag}::faulty_force_generalized"], atol=1e-15)
    assert np.array_equal(fs.ub, ref[f"{tag}::u_ub_physical"])                               # sys_model.py:240
    x, u = ref[f"{tag}::plant_x"], ref[f"{tag}::plant_u"]
    nxt = oracle.plant_rk4(x, u, fs, 0.1)                                                    # sys_model.py:138-226
    assert np.allclose(nxt, ref[f"{tag}::plant_next"], rtol=1e-13, atol=1e-14)
    assert np.allclose(oracle.normalize_quaternion_robot(nxt), ref[f"{tag}::plant_next_normalized"], rtol=1e-13, atol=1e-14)
    if tag == "nofault":
        return
    assert np.allclose(oracle.R_VEC, ref[f"{tag}::r"], atol=1e-16) and np.allclose(oracle.OMEGA_DES, ref[f"{tag}::omega_des"])
    assert np.allclose(oracle.F_VIRT, ref[f"{tag}::f_virt"][:3]) and np.allclose(fs.u_comp, ref[f"{tag}::compensation_force"], atol=1e-15)
    A, b = oracle.input_bounds(fs)                                                           # input_bounds.py:43-76, same row order
    assert A.shape == ref[f"{tag}::hull_A"].shape
    assert np.allclose(A, ref[f"{tag}::hull_A"], atol=1e-12) and np.allclose(b, ref[f"{tag}::hull_b"], atol=1e-12)
    c = oracle.robot_to_center(x)                                                            # spiral_model.py:91-109
    assert np.allclose(c, ref[f"{tag}::center"], rtol=1e-14, atol=1e-15)
    cn = oracle.spiral_rk4(c, ref[f"{tag}::spiral_u"], fs.generalized, 0.1)                  # spiral_model.py:44-76 + RK4
    assert np.allclose(cn, ref[f"{tag}::spiral_next"], rtol=1e-13, atol=1e-14)


def test_host_mirror_matches_reference(ft, ref):
    from ft_mpc_b200.controllers.tools.input_bounds import InputBounds
    from ft_mpc_b200.models import SpiralModel, SystemModel
    from ft_mpc_b200.util import BrokenThruster
    from ft_mpc_b200.util.get_trajectory import load_trajectory
    for tag, faults in TAGS.items():
        m = SystemModel(0.1)
        for i, a in faults:
            m.set_fault(BrokenThruster(i, a))
        assert np.array_equal(m.D, ref[f"{tag}::D"]) and np.array_equal(m.faulty_force, ref[f"{tag}::faulty_force"])
        assert np.array_equal(m.u_ub_physical, ref[f"{tag}::u_ub_physical"])
        if tag == "nofault":
            continue
        s = SpiralModel.from_system_model(m)
        assert np.allclose(s.r, ref[f"{tag}::r"], atol=1e-16)
        assert np.allclose(s.spiral_params.compensation_force, ref[f"{tag}::compensation_force"], atol=1e-15)
        A, b = InputBounds(m).get_conv_hull()
        assert np.allclose(A, ref[f"{tag}::hull_A"], atol=1e-12) and np.allclose(b, ref[f"{tag}::hull_b"], atol=1e-12)
        for k in range(3):
            assert np.allclose(s.robot_to_center(ref[f"{tag}::plant_x"][k]), ref[f"{tag}::center"][k], rtol=1e-14, atol=1e-15)
    for cmd in ("hover", "hover_1_-2_0.5", "generate_line", "generate_circle"):
        assert np.allclose(load_trajectory(cmd, 0.1, 3), ref[f"traj::{cmd}"], atol=1e-14), cmd


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["default", "single3dead"])
def test_cuda_kernels_match_reference(ft, built, ref, tag):
    """ftmpc_plant_step / ftmpc_robot_to_center / ftmpc_rk4_jac against the reference's own RK4 and transforms"""
    import torch
    from ft_mpc_b200 import _lib as L
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
    from ft_mpc_b200.models import SystemModel
    eng = BatchedMPC(SystemModel(0.1), 1, DEFAULT_Q, DEFAULT_R, [TAGS[tag]])
    d = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    x, u = d(ref[f"{tag}::plant_x"]), d(ref[f"{tag}::plant_u"])
    K = x.shape[0]
    raw = eng.plant_step(x, u, normalize=False).cpu().numpy()
    assert np.allclose(raw, ref[f"{tag}::plant_next"], rtol=1e-13, atol=1e-14)                 # model.dynamics(x, u), sim_env.py:85
    nrm = eng.plant_step(x, u, normalize=True).cpu().numpy()
    assert np.allclose(nrm, ref[f"{tag}::plant_next_normalized"], rtol=1e-13, atol=1e-14)      # sim_env.py:93
    c = torch.empty(K, 13, dtype=torch.float64, device="cuda")
    L.check(L.lib().ftmpc_robot_to_center(eng.handle, K, p(x), p(c), None))
    assert np.allclose(c.cpu().numpy(), ref[f"{tag}::center"], rtol=1e-14, atol=1e-15)
    # spiral RK4: the kernel takes the TOTAL body wrench u + D f_fault (spiral_model.py:61)
    wrench = d((ref[f"{tag}::spiral_u"] + ref[f"{tag}::faulty_force_generalized"])[:, None, :])
    xs = torch.zeros(K, 2, 13, dtype=torch.float64, device="cuda")
    xs[:, 0] = d(ref[f"{tag}::center"])
    jac = torch.zeros(K, 1, 13, 13, dtype=torch.float64, device="cuda")
    L.check(L.lib().ftmpc_rk4_jac(eng.handle, K, p(xs), p(wrench), p(jac), None, None, None))
    assert np.allclose(xs[:, 1].cpu().numpy(), ref[f"{tag}::spiral_next"], rtol=1e-13, atol=1e-14)


ASSIGN_CMDS = ("hover", "generate_line", "generate_circle", "circle_r_1.5_sPerFullCircle_12")


def test_assign_trajectory_matches_reference(ft, oracle, ref):
    """Row f-3: assign_trajectory (padding, omega_des rows, finite-difference nominal wrench) and the window slicing of
    the reference's controller (spiraling_mpc.py:255-286, 356-365) -- the reference methods were executed unbound on a
    namespace by tools/gen_ref_fixtures.py; the host mirror is run the same way (its constructor needs a GPU)."""
    import types
    from ft_mpc_b200.controllers import SpiralingController
    from ft_mpc_b200.util.get_trajectory import load_trajectory
    for cmd in ASSIGN_CMDS:
        ns = types.SimpleNamespace(Nt=20, dt=0.1, mass=16.8, device="cpu",
                                   spiral_params=types.SimpleNamespace(omega_des=np.array([0.0, 0.0, 0.6])))
        SpiralingController.assign_trajectory(ns, load_trajectory(cmd, 0.1, 3))
        assert np.allclose(ns.trajectory, ref[f"assign::{cmd}::trajectory"], rtol=0, atol=1e-14), cmd
        assert np.allclose(ns.nominal_input, ref[f"assign::{cmd}::nominal_input"], rtol=1e-12, atol=1e-11), cmd
        xr, ur = SpiralingController.get_next_trajectory_part(ns, 1.7)
        assert np.allclose(xr, ref[f"assign::{cmd}::window_x"], atol=1e-14) and np.allclose(ur, ref[f"assign::{cmd}::window_u"], rtol=1e-12, atol=1e-11)
        # (the padded tail of the line reference brakes, so only `hover` has an identically zero nominal wrench)
        assert ns._accelerating == bool(np.abs(ref[f"assign::{cmd}::nominal_input"]).max() > 0) and ns._accelerating == (cmd != "hover")
        # device tables are the transposed host tables; windows are contiguous [batch, N+1, .]
        ns.reference_window = types.MethodType(SpiralingController.reference_window, ns)
        w = SpiralingController.nominal_window(ns, 17, batch=3)
        if ns._accelerating:
            assert w.shape == (3, 21, 6) and w.is_contiguous() and np.allclose(w[1].numpy().T, ur, rtol=1e-12, atol=1e-11)
        else:
            assert w is None
        # the oracle's restatement agrees as well
        traj, nom = oracle.assign_trajectory(load_trajectory(cmd, 0.1, 3), 20, 0.1)
        assert np.allclose(traj, ref[f"assign::{cmd}::trajectory"], atol=1e-14) and np.allclose(nom, ref[f"assign::{cmd}::nominal_input"], rtol=1e-12, atol=1e-11)


def test_debug_export_matches_reference(ft, ref, tmp_path):
    """Row f-4: the 67-column ';'-CSV of ControllerDebug.export and the error definitions of DebugVal
    (controller_debug.py:9-79, 216-260).  The fixture is the file the reference's own classes wrote for a synthetic
    5-step history; the mirror must write the same bytes."""
    import types
    from ft_mpc_b200.util.controller_debug import HEADER, ControllerDebug
    holder = types.SimpleNamespace(model=types.SimpleNamespace(D=ref["debug::D"], faulty_force=np.zeros(16)))
    dbg = ControllerDebug.from_batch(holder, 0.1 * np.arange(5), ref["debug::states"], ref["debug::centers"],
                                     ref["debug::thrusts"], ref["debug::desired"])
    dbg.export(tmp_path / "dbg")
    got = (tmp_path / "dbg.csv").read_bytes()
    want = ref["debug::csv"].tobytes()
    assert len(HEADER) == 67 and dbg.table().shape == (5, 67)
    assert got == want
