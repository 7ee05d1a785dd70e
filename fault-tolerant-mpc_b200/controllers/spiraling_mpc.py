"""SpiralingController -- drop-in for ft_mpc/controllers/spiraling_mpc.py:23-365 backed by libftmpc.so.

Reference-facing surface kept (same names / argument meaning):
    SpiralingController(model, params, debug=None)        spiraling_mpc.py:27-44
    .load_trajectory(cmd, duration)                       :240-286
    .get_control(x0, t) -> ndarray[16]                    :288-317   (B = 1, stateful warm start)
and the batched entry the north star asks for:
    SpiralingController(model, horizon=.., weights={"Q":..,"R":..}, fault_mask=.. | fault_sets=[..])
    .step(state[B,13], ref[B,N+1,9], uref=None, scenario=None, warm=False) -> thrust[B,16] (torch, cuda)

Python here only prepares inputs (reference window, per-fault hull table) and owns the device tensors;
all per-step numerics run in CUDA behind include/ftmpc.h.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ..util.broken_thruster import BrokenThruster
from ..util.controller_debug import DebugVal
from ..util.get_trajectory import load_trajectory
from .tools.input_bounds import hull_of_faults
from .tools.spiral_parameters import SpiralParameters

DEFAULT_Q = [1, 1, 1, 1, 1, 1, 2, 2, 2]                  # ft_mpc/config/reactive.yaml:32
DEFAULT_R = [0.1, 0.1, 0.1, 0.01, 0.01, 0.01]           # ft_mpc/config/reactive.yaml:33


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("ft_mpc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if dev.type != "cuda":
        raise RuntimeError("ft_mpc_b200 tensors must live on a CUDA device")
    return dev


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def hull_table_entry(A, b) -> np.ndarray:
    """One row of the hull table: A_h padded to 26x6 with zero rows, b_h padded with 1e30 (never active)."""
    e = np.zeros(L.HULL_STRIDE)
    Ap, bp = np.zeros((L.NH, L.NU)), np.full(L.NH, 1e30)
    if len(b) > L.NH:
        raise ValueError(f"hull has {len(b)} facets, table rows hold {L.NH}")
    Ap[:len(b)], bp[:len(b)] = A, b
    e[:L.NH * L.NU], e[L.NH * L.NU:] = Ap.ravel(), bp
    return e


class BatchedMPC:
    """Thin owner of one ftmpc handle + the device tensors of a batch.  `fault_sets` is a list of fault
    scenarios, each a list of (thruster index, intensity); instance i uses scenario[i]."""

    def __init__(self, model, horizon, Q, R, fault_sets, device=None, **solver_opts):
        self.device = _require_cuda(device)
        self.lib = L.lib()
        self.N = int(horizon)
        self.model = model
        sp = SpiralParameters(model)
        self.omega_des, self.f_virt, self.r = sp.omega_des, sp.f_virt, sp.r
        # a scenario is a list of (index, intensity) or a dict {faults, A, b} carrying a precomputed hull
        self.fault_sets = [list(fs["faults"]) if isinstance(fs, dict) else list(fs) for fs in fault_sets]
        table, self.fault_force_tab, self.mask_tab = [], [], []
        for fs, src in zip(self.fault_sets, fault_sets):
            if isinstance(src, dict):
                A, b = src["A"], src["b"]
            else:
                A, b = hull_of_faults(model.D, model.max_thrust, fs)       # input_bounds.py:43-76
            table.append(hull_table_entry(A, b))
            ff = np.zeros(L.NTHR)
            m = 0
            for i, inten in fs:
                ff[i] = inten * model.max_thrust                           # sys_model.py:239
                m |= 1 << i
            self.fault_force_tab.append(ff)
            self.mask_tab.append(m)
        self.hull_table = np.ascontiguousarray(np.stack(table))
        self.cfg = L.make_config(self.N, Q, R, dt=model.dt, mass=model.mass, inertia=model.inertia, r=self.r,
                                 f_virt=self.f_virt, max_thrust=model.max_thrust, D=model.D,
                                 n_hull_sets=len(self.fault_sets), **solver_opts)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.ftmpc_create(C.byref(self.handle), C.byref(self.cfg),
                                          self.hull_table.ctypes.data_as(C.POINTER(C.c_double))), "ftmpc_create")
        self.nz = self.lib.ftmpc_num_var(self.handle)
        self.mc = self.lib.ftmpc_num_ineq(self.handle)
        self._fault_force_dev = torch.tensor(np.stack(self.fault_force_tab), dtype=torch.float64, device=self.device)
        self._mask_dev = torch.tensor(np.array(self.mask_tab, dtype=np.int64), device=self.device)
        self._bufs = {}
        self._scen0 = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.ftmpc_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def buffers(self, B, slot=0):
        """Output / workspace tensors for a batch of B (cached).  `slot` selects an independent set: steps that use
        different sets may be in flight on different CUDA streams at the same time (the work queue of a launch lives in
        its workspace), which lets the tail of one batch overlap the head of the next."""
        b = self._bufs.get((B, slot))
        if b is None:
            nbytes = C.c_size_t()
            L.check(self.lib.ftmpc_workspace_bytes(self.handle, B, C.byref(nbytes)), "ftmpc_workspace_bytes")
            dev, f64 = self.device, torch.float64
            b = dict(ws=torch.empty(nbytes.value, dtype=torch.uint8, device=dev),
                     z=torch.zeros(B, self.nz, dtype=f64, device=dev),
                     thrust=torch.empty(B, L.NTHR, dtype=f64, device=dev),
                     u0=torch.empty(B, L.NU, dtype=f64, device=dev),
                     active=torch.empty(B, (self.mc + 31) // 32, dtype=torch.int32, device=dev),
                     status=torch.empty(B, dtype=torch.int32, device=dev),
                     iters=torch.empty(B, 2, dtype=torch.int32, device=dev),
                     cost=torch.empty(B, dtype=f64, device=dev))
            self._bufs[(B, slot)] = b
        return b

    def _default_scenario(self, B):
        """(mask, fault_force, hull_idx) of scenario 0 for every instance, built once per batch size"""
        sc = self._scen0.get(B)
        if sc is None:
            sc = self.scenario_tensors(torch.zeros(B, dtype=torch.int64, device=self.device))
            self._scen0[B] = sc
        return sc

    def scenario_tensors(self, scenario):
        """scenario [B] int (index into fault_sets) -> (mask uint16 [B], fault_force [B,16], hull_idx int32 [B])"""
        scenario = scenario.to(self.device, torch.int64)
        mask = self._mask_dev[scenario].to(torch.int16)            # bit pattern reinterpreted as uint16 by the library
        return mask.contiguous(), self._fault_force_dev[scenario].contiguous(), scenario.to(torch.int32).contiguous()

    def step(self, state, xref, uref=None, scenario=None, warm=False, out=None):
        """All arguments are CUDA fp64 tensors; returns the dict of output tensors (no synchronisation)."""
        B = state.shape[0]
        b = out or self.buffers(B)
        if scenario is None:
            scenario = self._default_scenario(B)
        mask, ff, hidx = scenario if isinstance(scenario, tuple) else self.scenario_tensors(scenario)
        assert state.shape == (B, L.NX) and xref.shape == (B, self.N + 1, L.NE)
        assert state.is_contiguous() and xref.is_contiguous() and state.dtype == torch.float64
        assert uref is None or (uref.shape == (B, self.N + 1, L.NU) and uref.is_contiguous() and uref.dtype == torch.float64)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.ftmpc_step(self.handle, B, _ptr(state), _ptr(xref), _ptr(uref), _ptr(mask), _ptr(ff),
                                    _ptr(hidx), int(bool(warm)), _ptr(b["z"]), _ptr(b["thrust"]), _ptr(b["u0"]),
                                    _ptr(b["active"]), _ptr(b["status"]), _ptr(b["iters"]), _ptr(b["cost"]),
                                    _ptr(b["ws"]), b["ws"].numel(), C.c_void_p(stream)), "ftmpc_step")
        return b

    def plant_step(self, state, thrust, scenario=None, noise=None, normalize=True):
        B = state.shape[0]
        if scenario is None:
            scenario = torch.zeros(B, dtype=torch.int64, device=self.device)
        mask, ff, _ = scenario if isinstance(scenario, tuple) else self.scenario_tensors(scenario)
        nxt = torch.empty_like(state)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.ftmpc_plant_step(self.handle, B, _ptr(state), _ptr(thrust), _ptr(mask), _ptr(ff), _ptr(noise),
                                          int(bool(normalize)), _ptr(nxt), C.c_void_p(stream)), "ftmpc_plant_step")
        return nxt

    def hull_facets(self, fault_sets):
        """Input-bound polytopes of arbitrary fault sets built on the device (ftmpc_hull_facets; analytic facet enumeration,
        SURVEY.md section 8 row f-2).  fault_sets: list of [(thruster, intensity), ...].  Returns (table [n, HULL_STRIDE],
        n_rows [n], status [n]) as CUDA tensors; the table rows are what ftmpc_create takes."""
        n = len(fault_sets)
        mask = np.zeros(n, np.int16)
        ff = np.zeros((n, L.NTHR))
        for k, fs in enumerate(fault_sets):
            m = 0
            for i, a in fs:
                m |= 1 << int(i)
                ff[k, int(i)] = float(a) * self.model.max_thrust
            mask[k] = np.array(m, dtype=np.uint16).view(np.int16)
        mask_d = torch.tensor(mask, device=self.device)
        ff_d = torch.tensor(ff, dtype=torch.float64, device=self.device)
        table = torch.empty(n, L.HULL_STRIDE, dtype=torch.float64, device=self.device)
        nrows = torch.empty(n, dtype=torch.int32, device=self.device)
        status = torch.empty(n, dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.ftmpc_hull_facets(self.handle, n, _ptr(mask_d), _ptr(ff_d), _ptr(table), _ptr(nrows), _ptr(status),
                                           C.c_void_p(stream)), "ftmpc_hull_facets")
        return table, nrows, status

    def closed_loop(self, state0, trajectory, scenario=None, steps=1, noise=None, start_step=0, nominal_input=None):
        """`steps` x (ftmpc_step -> ftmpc_plant_step) entirely on the device: SimulationEnvironment.run_simulation
        (sim_env.py:77-112) for a batch.  trajectory: [T,9] device reference table (assign_trajectory), shared by
        all instances; nominal_input: optional [T,6] device table of the nominal wrench of an accelerating reference
        (assign_trajectory, spiraling_mpc.py:279-286); noise: optional [steps,B,13] tensor added after each plant step
        (the reference draws unseeded U(0,1e-3), sim_env.py:88-91).  Warm start from the second step on (spiraling_mpc.py:324-331).
        Returns (final_state [B,13], cumulative optimal cost [B], worst status [B], steps done [B])."""
        B = state0.shape[0]
        sc = self._default_scenario(B) if scenario is None else (scenario if isinstance(scenario, tuple) else self.scenario_tensors(scenario))
        mask, ff, hidx = sc
        b = self.buffers(B)
        state = state0.clone().contiguous()
        trajectory = trajectory.contiguous()
        assert trajectory.dim() == 2 and trajectory.shape[1] == L.NE and trajectory.dtype == torch.float64
        if nominal_input is not None:
            nominal_input = nominal_input.contiguous()
            assert nominal_input.shape == (trajectory.shape[0], L.NU)
        if noise is not None:
            noise = noise.contiguous()
            assert noise.shape[0] >= steps and noise.shape[1:] == (B, L.NX)
        cost = torch.empty(B, dtype=torch.float64, device=self.device)
        worst = torch.empty(B, dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        # one native call enqueues all steps (3 launches each) on the stream: no Python in the loop, no per-step window copies
        L.check(self.lib.ftmpc_closed_loop(self.handle, B, int(steps), int(start_step), int(trajectory.shape[0]), _ptr(state),
                                           _ptr(trajectory), _ptr(nominal_input), _ptr(mask), _ptr(ff), _ptr(hidx), _ptr(noise), 0,
                                           _ptr(b["z"]), _ptr(b["thrust"]), _ptr(b["u0"]), _ptr(b["active"]), _ptr(b["status"]),
                                           _ptr(b["iters"]), _ptr(b["cost"]), _ptr(cost), _ptr(worst), _ptr(b["ws"]),
                                           b["ws"].numel(), C.c_void_p(stream)), "ftmpc_closed_loop")
        done = torch.full((B,), steps, dtype=torch.int32, device=self.device)
        return state, cost, worst, done


class _PlantStepper:
    """model.dynamics(x, u) for the host-facing SystemModel (sim_env.py:85): one un-normalised RK4 step."""

    def __init__(self, model):
        faults = [(bt.index, bt.intensity) for bt in model.broken_thrusters]
        # the plant step reads only the model constants and the fault record of the handle: a placeholder hull entry keeps
        # the construction free of the Qhull pass (input_bounds.py:43-76) a controller would need
        self.eng = BatchedMPC(model, 1, DEFAULT_Q, DEFAULT_R, [{"faults": faults, "A": np.zeros((1, L.NU)), "b": np.ones(1)}])

    def __call__(self, x, u):
        dev = self.eng.device
        xs = torch.tensor(x, dtype=torch.float64, device=dev)
        us = torch.tensor(u, dtype=torch.float64, device=dev)
        return self.eng.plant_step(xs, us, normalize=False).cpu().numpy()


class SpiralingController:
    """Controller implementing micro-orbiting (reference: spiraling_mpc.py:23)."""

    def __init__(self, model, params=None, debug=None, *, horizon=None, weights=None, fault_mask=None,
                 fault_intensity=None, fault_sets=None, device=None, **solver_opts):
        self.model = model
        self.debug = debug
        params = dict(params or {})
        if horizon is not None:
            params["horizon"] = int(horizon)
        if weights is not None:
            params["param_set"] = "P1"
            params["P1"] = {"Q": list(weights["Q"]), "R": list(weights["R"])}
        params.setdefault("horizon", 15)                                  # reactive.yaml:26
        params.setdefault("param_set", "P1")
        params.setdefault(params["param_set"], {"Q": DEFAULT_Q, "R": DEFAULT_R})
        if params.get("xub") is not None or params.get("xlb") is not None:
            raise NotImplementedError("state bounds params['xub'] / params['xlb'] (spiraling_mpc.py:129-130, 180-185) are refused: "
                                      "the reference's configuration never sets them and the reduced-space SQP has no "
                                      "state-bound rows (DESIGN.md section 7)")
        self.params = params
        self.Nt = params["horizon"]
        ps = params[params["param_set"]]
        self.Q, self.R = np.diag(ps["Q"]), np.diag(ps["R"])               # spiraling_mpc.py:91-93
        self.spiral_params = SpiralParameters(model)
        self.mass, self.J, self.dt = model.mass, model.inertia, model.dt
        self.Nx, self.Nu, self.Nopt = 13, 6, 9
        if fault_sets is None:
            if fault_mask is not None:
                inten = fault_intensity if fault_intensity is not None else [0.0] * 16
                fault_sets = [[(i, float(inten[i])) for i in range(16) if (int(fault_mask) >> i) & 1]]
            else:
                fault_sets = [[(bt.index, bt.intensity) for bt in model.broken_thrusters]]
        # params["solver_opts"] carries IPOPT / nlpsol options in the reference (spiraling_mpc.py:227-229, e.g.
        # "ipopt.max_iter"); they have no meaning for this solver and are ignored with a note, except the ones that map:
        # ipopt.max_iter -> max_sqp_iter.  Native options go in params["ftmpc_opts"] (or as keyword arguments).
        opts = {}
        for key, val in dict(params.get("solver_opts") or {}).items():
            if key == "ipopt.max_iter":
                opts["max_sqp_iter"] = int(val)
            elif key in L.SOLVER_OPTION_NAMES:
                opts[key] = val
            else:
                self.ignored_solver_opts = getattr(self, "ignored_solver_opts", []) + [key]
        opts.update(dict(params.get("ftmpc_opts") or {}))
        opts.update(solver_opts)
        self.engine = BatchedMPC(model, self.Nt, ps["Q"], ps["R"], fault_sets, device=device, **opts)
        self.device = self.engine.device
        self.trajectory = None
        self.nominal_input = None
        self.optimal_solution = None
        self.last_status = None

    # ---- reference trajectory ------------------------------------------------- spiraling_mpc.py:240-286
    def load_trajectory(self, cmd, duration):
        self.assign_trajectory(load_trajectory(cmd, self.dt, duration))                  # spiraling_mpc.py:252

    def assign_trajectory(self, trajectory):
        N = self.Nt
        orig = np.hstack((trajectory, np.tile(trajectory[:, -1:], (1, N))))                     # :264
        om = np.tile(self.spiral_params.omega_des, (orig.shape[1], 1)).T                        # :269-272
        self.trajectory = np.concatenate((orig[0:6, :], om))                                    # :274-277
        second = np.gradient(np.gradient(self.trajectory[0:3, :], axis=1), axis=1) / self.dt ** 2   # :283
        self.nominal_input = np.vstack((second * self.mass, np.zeros_like(second)))             # :285
        self._traj_dev = torch.tensor(self.trajectory.T.copy(), dtype=torch.float64, device=self.device)
        # accelerating references: the nominal wrench travels with the window and is rotated into the body frame per
        # stage inside the kernels (stage_wrench, spiraling_mpc.py:156-166); a hover reference keeps the NULL fast path
        self._accelerating = bool(np.abs(self.nominal_input).max() > 0.0)
        self._uref_dev = torch.tensor(self.nominal_input.T.copy(), dtype=torch.float64, device=self.device)

    def get_next_trajectory_part(self, t):
        """Window of N+1 reference points starting at int(t/dt) (spiraling_mpc.py:356-365)."""
        i = int(t / self.dt)
        return self.trajectory[:, i:i + self.Nt + 1], self.nominal_input[:, i:i + self.Nt + 1]

    def reference_window(self, step_index, batch=1):
        """[batch, N+1, 9] device window at an integer step index (avoids the int(t/dt) rounding quirk)."""
        w = self._traj_dev[step_index:step_index + self.Nt + 1]
        return w.unsqueeze(0).expand(batch, -1, -1).contiguous()

    def nominal_window(self, step_index, batch=1):
        """[batch, N+1, 6] device window of the nominal wrench, or None for a reference without acceleration."""
        if not self._accelerating:
            return None
        w = self._uref_dev[step_index:step_index + self.Nt + 1]
        return w.unsqueeze(0).expand(batch, -1, -1).contiguous()

    # ---- single-instance reference API ------------------------------------------ spiraling_mpc.py:288-317
    def _single_io(self):
        """Staging of the B = 1 call: a pinned host state, and ONE device record [thrust 16 | u0 6 | cost 1 | status, sqp
        iterations, qp iterations (int32) + pad] that the kernels write in place and that comes back in one pinned copy."""
        io = getattr(self, "_io1", None)
        if io is None:
            dev = self.device
            rec = torch.zeros(25, dtype=torch.float64, device=dev)          # 23 doubles + 4 int32
            ints = rec[23:25].view(torch.int32)
            eng = self.engine
            nbytes = C.c_size_t()
            L.check(eng.lib.ftmpc_workspace_bytes(eng.handle, 1, C.byref(nbytes)), "ftmpc_workspace_bytes")
            out = dict(ws=torch.empty(nbytes.value, dtype=torch.uint8, device=dev),
                       z=torch.zeros(1, eng.nz, dtype=torch.float64, device=dev),
                       thrust=rec[0:16].view(1, 16), u0=rec[16:22].view(1, 6), cost=rec[22:23],
                       status=ints[0:1], iters=ints[1:3].view(1, 2),
                       active=torch.empty(1, (eng.mc + 31) // 32, dtype=torch.int32, device=dev))
            io = dict(out=out, rec=rec, rec_host=torch.zeros(25, dtype=torch.float64).pin_memory(),
                      state_host=torch.zeros(1, 13, dtype=torch.float64).pin_memory(),
                      state=torch.zeros(1, 13, dtype=torch.float64, device=dev))
            self._io1 = io
        return io

    def get_control(self, x0, t):
        io = self._single_io()
        x0 = np.asarray(x0, float).reshape(13)
        io["state_host"].numpy()[0] = x0
        io["state"].copy_(io["state_host"], non_blocking=True)
        k = int(t / self.dt)                                                               # spiraling_mpc.py:360
        # the window of a single instance is a contiguous slice of the device table: no copy
        xref = self._traj_dev[k:k + self.Nt + 1].unsqueeze(0)
        uref = self._uref_dev[k:k + self.Nt + 1].unsqueeze(0) if self._accelerating else None
        out = self.engine.step(io["state"], xref, uref, warm=self.optimal_solution is not None, out=io["out"])
        io["rec_host"].copy_(io["rec"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()                               # the one host round trip
        self.optimal_solution = out["z"]
        rec = io["rec_host"].numpy()
        ints = rec[23:25].view(np.int32)
        self.last_status = int(ints[0])
        self.last_iters = (int(ints[1]), int(ints[2]))
        self.last_u0 = rec[16:22].copy()
        self.last_cost = float(rec[22])
        thrust = rec[0:16].copy()
        if self.debug is not None:                                                         # spiraling_mpc.py:309-315
            dv = DebugVal(self, t)
            dv.set_state(x0)
            dv.set_circle_state(out["z"][0, 6 * self.Nt:6 * self.Nt + 9].cpu().numpy())     # c0 = x_0 of the decision vector
            dv.set_input(thrust, self.model)
            dv.set_desired_state(self.trajectory[:, k])
            dv.calculate_errors()
            self.debug.add_debug_val(dv)
        return thrust

    # ---- batched API -------------------------------------------------------------------------------------
    def step(self, state, ref, uref=None, scenario=None, warm=False, slot=0):
        """state [B,13] robot states, ref [B,N+1,9] reference windows -> thrust [B,16] (CUDA tensor).
        Other outputs of the solve are in `self.last` (u0, active, status, iters, cost, z).  Work is enqueued on the
        current CUDA stream; steps issued with different `slot`s use independent output / workspace buffers and may
        therefore overlap on different streams (a warm start continues the `z` of its own slot)."""
        state = torch.as_tensor(state, dtype=torch.float64, device=self.device).contiguous()
        ref = torch.as_tensor(ref, dtype=torch.float64, device=self.device).contiguous()
        self.last = self.engine.step(state, ref, uref, scenario, warm, out=self.engine.buffers(state.shape[0], slot))
        return self.last["thrust"]


__all__ = ["SpiralingController", "BatchedMPC", "BrokenThruster", "hull_table_entry"]
