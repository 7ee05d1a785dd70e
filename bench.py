#!/usr/bin/env python
"""bench.py -- batched MPC solves/s of the B200-native ft_mpc drop-in (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B_per_gpu] [--horizon 20] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the 1e6 solves/s target is quoted on; weak scaling: the per-GPU
shard is fixed at 65,536 / 8 = 8,192 instances): Monte-Carlo fault scenarios -- all well-posed single and
double thruster failures (dead / stuck-on), seeded random initial robot states, hover reference, horizon
N = 20, cold start.  One "step" = one ftmpc_step over the rank's shard = B complete get_control equivalents
(state -> SQP on the reference NLP -> 16 thrusts).  Only converged instances (status 0) count as solves.
Consecutive steps are independent batches; they are enqueued on `--streams` (default 4) alternating CUDA streams, each
with its own output / workspace buffers, so the SMs that run out of instances at the tail of one launch start on the next
batch (`--streams 1`: strictly serial steps).  The timing events sit on the main stream, which all side streams wait
for at the start and which waits for all of them at the end.

One JSON line on stdout (rank 0).  `value` is device-resident throughput, `e2e` goes through the public
SpiralingController.step with pinned HOST buffers (H2D of the inputs and D2H of thrust/status inside the timed
region), `roofline` compares the dominant kernel -- timed in one more step run alone, with the in-kernel phase profile
on -- with the measured FP64 FMA peak of the device (the solve is FP64-compute/latency bound) and reports
the HBM side too, `cpu_baseline` / `--impl reference`
time the same algorithm on the host cores (oracle/cpu_port -- the reference's own casadi/IPOPT stack is not
installable offline).  Extra keys at N = 1 GPU: `latency` (BASELINE configs[1]: single-instance warm-started
get_control over the 300-step demo loop, p50 / p95) and `batch_1024` (configs[2] size: 1024 instances per step, four
independent batches in flight -- a launch of 1024 ends on its slowest instance, so small batches are pipelined deeper).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "batched MPC solves/sec (fault-scenario get_control equivalents, N=20)"
UNIT = "solves/s"


# ---------------------------------------------------------------------------------------------------------
# workload + flop/byte model
# ---------------------------------------------------------------------------------------------------------
def make_workload(batch_total: int, N: int, seed: int = 1):
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200.util import scenarios
    cells = scenarios.load_cells(kinds=("single", "double"))
    states = scenarios.random_states(batch_total, seed)
    scen = (np.arange(batch_total) * 7919) % len(cells)          # every cell appears, decorrelated from the state index
    xref = scenarios.hover_reference(batch_total, N)
    return cells, states, scen, xref


def algorithmic_flops(N: int, k_sqp: np.ndarray, k_qp: np.ndarray) -> dict:
    """F = K_sqp (F_lin + F_cond + F_chol) + K_qp F_iter + F_alloc, per instance."""
    n, m = 6 * N, 26 * N + 72
    f_lin = N * 29000.0
    f_cond = N * (N + 1) / 2 * (2 * 169 * 6) + N * N / 2 * (2 * 169 * 6 + 2 * 13 * 36) + 288.0 * n
    f_chol = n ** 3 / 3.0
    f_iter = 2.0 * n * n + 624.0 * N + 36.0 * n + 10.0 * (n + m)
    f_alloc = 5000.0
    ks, kq = float(k_sqp.sum()), float(k_qp.sum())
    return dict(total=ks * (f_lin + f_cond + f_chol) + kq * f_iter + f_alloc * len(k_sqp),
                qp_kernel=ks * (f_cond + f_chol) + kq * f_iter, lin_kernel=ks * f_lin,
                per_sqp_iter=f_lin + f_cond + f_chol, per_qp_iter=f_iter)


def algorithmic_bytes(N: int) -> float:
    """compulsory HBM I/O per solve: inputs + warm start in/out + outputs."""
    nz, m = 6 * N + 13 * (N + 1), 26 * N + 72
    return 8.0 * (13 + 9 * (N + 1) + 6 * (N + 1) + nz) + 8 + 8.0 * (nz + 16 + 6) + 4.0 * ((m + 31) // 32) + 12


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML in-process, `nvidia-smi` as fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run_nvml(self) -> bool:
        """in-process NVML sampling (nvidia_ml_py): no process spawn, no NVML re-initialisation per sample -- an `nvidia-smi`
        child every 200 ms perturbs the timed region it is supposed to observe (measured: device-resident steps 3-15 %
        slower than the later, unobserved e2e steps).  Returns False when NVML is not usable (falls back to nvidia-smi)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        bits = {0x8: 2, 0x40: 3, 0x20: 4, 0x4: 5}      # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap -> row slot
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "Not Active", "Not Active", "Not Active", "Not Active"]
                r = int(get_reasons(h))
                for b, slot in bits.items():
                    if r & b:
                        row[slot] = "Active"
                self.rows.append(row)
            except Exception:
                pass
            self._stop.wait(0.1)
        try:
            nv.nvmlShutdown()
        except Exception:
            pass
        return True

    def _run(self):
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(6)

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm (reference arm and cpu_baseline): the same NLP solved on the host cores by oracle/cpu_port
# ---------------------------------------------------------------------------------------------------------
def cpu_info() -> dict:
    model = None
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    flags = None
    try:
        for ln in open(ROOT / "oracle" / "Makefile"):
            if "$(CXX)" in ln and "-O" in ln:
                flags = " ".join(w for w in ln.split() if w.startswith("-") and not w.startswith("-I") and w not in ("-x", "-o"))
    except OSError:
        pass
    return {"cpu_model": model, "compiler_flags": flags, "host_cores": os.cpu_count()}


def cpu_solve_rate(N: int, cells, states, scen, xref, sample: int, steps: int, warmup: int):
    """Times `steps` passes over the first `sample` instances with all host threads.  Returns (solves/s, ms/step, cores, ok_frac)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import helpers as H                      # ctypes binding of oracle/_cpu/libftmpc_cpu.so (checker code)
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200 import _lib as L
    from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, hull_table_entry
    from ft_mpc_b200.controllers.tools.spiral_parameters import SpiralParameters
    from ft_mpc_b200.models import SystemModel
    if not H.CPU_LIB.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    port = H.CpuPort()
    model = SystemModel(0.1)
    sp = SpiralParameters(model)
    table = np.ascontiguousarray(np.stack([hull_table_entry(c["A"], c["b"]) for c in cells]))
    cfg = L.make_config(N, DEFAULT_Q, DEFAULT_R, dt=model.dt, mass=model.mass, inertia=model.inertia, r=sp.r, f_virt=sp.f_virt,
                        max_thrust=model.max_thrust, D=model.D, n_hull_sets=len(cells))
    masks = np.zeros(len(cells), np.uint16)
    ffs = np.zeros((len(cells), 16))
    for k, c in enumerate(cells):
        for i, a in c["faults"]:
            masks[k] |= np.uint16(1 << i)
            ffs[k, i] = a * model.max_thrust
    s = slice(0, sample)
    args = (cfg, table, states[s], xref[s], None, masks[scen[s]], ffs[scen[s]], scen[s])
    cores = os.cpu_count() or port.lib.ftmpc_cpu_num_threads()       # torchrun exports OMP_NUM_THREADS=1: ask for all cores explicitly
    args = args + (0, None, cores)
    for _ in range(warmup):
        port.step(*args)
    t0 = time.perf_counter()
    ok = 0
    for _ in range(steps):
        out = port.step(*args)
        ok += int((out["status"] == 0).sum())
    dt = time.perf_counter() - t0
    return ok / dt, dt / steps * 1e3, cores, ok / (steps * sample)


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)           # a multiple of the number of streams
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8192, help="instances per GPU (weak scaling)")
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances per CPU step (0 = auto)")
    ap.add_argument("--streams", type=int, default=4, help="independent batches in flight (1 = strictly serial steps)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--solver-opts", default="{}", help='JSON dict of ftmpc_config overrides for experiments, e.g. {"warm_qp": 1}')
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = a.horizon
    W = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    config = {"workload": f"mc_fault_scenarios single+double(dead|stuck) well-posed cells, N={N}, cold start, hover, "
                          f"{a.batch} instances per GPU (BASELINE configs[3] shard)", "horizon": N, "batch_per_gpu": a.batch,
              "global_batch": a.batch * world, "seed": 1, "parallelism": f"dp{world} (independent instances, no data-path collective)",
              "l2_policy": "per-step working set (inputs + per-instance workspace) exceeds the 126 MB L2",
              "pipelining": (f"{a.streams} CUDA streams: consecutive steps (independent batches, separate output/workspace buffers) are enqueued on "
                             "alternating streams, so the SMs that run out of instances at the tail of one launch start on the next batch"
                             if a.streams > 1 else "none (steps strictly serial on one stream)")}
    if a.solver_opts != "{}":
        config["solver_opts"] = json.loads(a.solver_opts)

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if a.impl == "reference":
        if rank != 0:
            return 0
        sample = min(a.cpu_sample or 2048, a.batch)
        cells, states, scen, xref = make_workload(a.batch * world, N)       # the batch the B200 arm solves (rank 0's shard first)
        rate, ms, cores, okf = cpu_solve_rate(N, cells, states, scen, xref, sample, a.steps, a.warmup)
        config["reference_sample_per_step"] = sample
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"the first {sample} instances of the batch the B200 arm solves (same seed, same order), every step, all "
                                           f"{cores} host threads (OpenMP, one instance per thread); "
                                           "the reference's casadi/IPOPT solve is not installable offline, this is the C++ port of the same SQP "
                                           "on the reference NLP (oracle/cpu_port)", "converged_frac": okf, **cpu_info()},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    import ftmpc_import
    ftmpc_import.load()
    from ft_mpc_b200 import _lib as L
    from ft_mpc_b200.controllers import SpiralingController
    from ft_mpc_b200.distributed import gather_results, pack_results, shard_bounds
    from ft_mpc_b200.models import SystemModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ft_mpc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    devs = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")                # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device(devs))
    B = a.batch
    Btot = B * world
    cells, states, scen, xref = make_workload(Btot, N)
    lo, hi = shard_bounds(Btot, rank, world)
    ctrl = SpiralingController(SystemModel(0.1), horizon=N, weights={"Q": [1, 1, 1, 1, 1, 1, 2, 2, 2], "R": [.1, .1, .1, .01, .01, .01]},
                               fault_sets=cells, device=devs, **json.loads(a.solver_opts))
    eng = ctrl.engine
    f64 = torch.float64
    st_d = torch.tensor(states[lo:hi], dtype=f64, device=devs)
    xr_d = torch.tensor(xref[lo:hi], dtype=f64, device=devs)
    sc_t = eng.scenario_tensors(torch.tensor(scen[lo:hi], device=devs))
    stream = torch.cuda.current_stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
        eng.step(st_d, xr_d, scenario=sc_t)
    sync_all()
    # ---- timed region: device-resident inputs.  Step i runs on stream i % S with buffer set i % S; the timing events sit
    # on the main stream, which every side stream waits for at the start and which waits for every side stream at the end.
    S = max(1, a.streams)
    side = [torch.cuda.Stream(device=devs) for _ in range(S)] if S > 1 else [stream]
    ok_warm = torch.zeros(S, dtype=torch.int64, device=devs)
    for i in range(2 * S):                                        # allocate the buffer sets outside the timed region and run the
        with torch.cuda.stream(side[i % S]):                      # pipelined pattern once untimed: first use of the side streams
            o_ = eng.step(st_d, xr_d, scenario=sc_t, out=eng.buffers(hi - lo, i % S))
            ok_warm[i % S] += (o_["status"] == 0).sum()           # ... and of the counting kernels (their lazy module load inside the
    sync_all()                                                    # timed region cost 60 ms once: tools/bench_valueloop_probe.py)

    def fan_out(ev):
        for s_ in side:
            if s_ is not stream:
                s_.wait_event(ev)

    def fan_in():
        for s_ in side:
            if s_ is not stream:
                ev = torch.cuda.Event()
                ev.record(s_)
                stream.wait_event(ev)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ok_acc = torch.zeros(S, dtype=torch.int64, device=devs)       # converged instances of EVERY timed step, per stream slot
    with ClockSampler(local) as clk:
        sync_all()
        e0.record(stream)
        fan_out(e0)
        launches = 0
        for i in range(a.steps):
            with torch.cuda.stream(side[i % S]):
                out = eng.step(st_d, xr_d, scenario=sc_t, out=eng.buffers(hi - lo, i % S))
                ok_acc[i % S] += (out["status"] == 0).sum()
            launches += eng.lib.ftmpc_last_launches(eng.handle)
        fan_in()
        e1.record(stream)
        sync_all()
    ms_total = e0.elapsed_time(e1)
    # one more step, alone and with the in-kernel profile switched on (per-handle counters and events, so not inside the
    # pipelined region): device time of the two kernels and the solver kernel's per-phase cycle profile
    L.check(eng.lib.ftmpc_profile_enable(eng.handle, 1))
    out = eng.step(st_d, xr_d, scenario=sc_t)
    kms = (C.c_double * 2)(); cyc = (C.c_int64 * L.N_PHASES)()
    L.check(eng.lib.ftmpc_profile_read(eng.handle, C.c_void_p(stream.cuda_stream), kms, cyc, L.N_PHASES))
    kernel_ms = {"k_solve": float(kms[0]), "k_alloc": float(kms[1])}
    cyc = np.array(list(cyc), dtype=float)
    L.check(eng.lib.ftmpc_profile_enable(eng.handle, 0))
    status = out["status"].cpu().numpy()
    iters = out["iters"].cpu().numpy()
    ok_local = int(ok_acc.sum().item())                           # summed over all timed steps
    t = torch.tensor([ms_total, float(ok_local)], dtype=f64, device=devs)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, ok_total = float(tmax[0]), float(tsum[1])
    else:
        ok_total = float(ok_local)
    value = ok_total / (ms_total * 1e-3)

    # ---- the one collective of the path: all-gather of the per-instance result records (outside the timed solve)
    rec = pack_results(st_d, out["cost"], out["status"], torch.ones(hi - lo, dtype=torch.int32, device=devs))
    full = gather_results(rec, Btot)
    assert full.shape[0] == Btot

    # ---- e2e: public API, pinned host buffers, H2D + D2H inside the timed region
    st_h = torch.tensor(states[lo:hi], dtype=f64).pin_memory()
    xr_h = torch.tensor(xref[lo:hi], dtype=f64).pin_memory()
    sc_h = torch.tensor(scen[lo:hi], dtype=torch.int64).pin_memory()
    th_h = [torch.empty(hi - lo, 16, dtype=f64).pin_memory() for _ in range(S)]
    stt_h = [torch.empty(hi - lo, dtype=torch.int32).pin_memory() for _ in range(S)]

    def e2e_step(slot):
        with torch.cuda.stream(side[slot]):
            thrust = ctrl.step(st_h.to(devs, non_blocking=True), xr_h.to(devs, non_blocking=True),
                               scenario=sc_h.to(devs, non_blocking=True), slot=slot)
            th_h[slot].copy_(thrust, non_blocking=True)
            stt_h[slot].copy_(ctrl.last["status"], non_blocking=True)
            ok_e2e_acc[slot] += (ctrl.last["status"] == 0).sum()

    e2e_steps = a.steps
    ok_e2e_acc = torch.zeros(S, dtype=torch.int64, device=devs)
    for i in range(S):
        e2e_step(i)
    sync_all()
    ok_e2e_acc.zero_()
    sync_all()
    e0.record(stream)
    fan_out(e0)
    for i in range(e2e_steps):
        e2e_step(i % S)
    fan_in()
    e1.record(stream)
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    ok_e2e = float(ok_e2e_acc.sum().item())                       # all timed steps
    te = torch.tensor([ms_e2e, ok_e2e], dtype=f64, device=devs)
    if world > 1:
        temax = te.clone(); dist.all_reduce(temax, op=dist.ReduceOp.MAX)
        tesum = te.clone(); dist.all_reduce(tesum, op=dist.ReduceOp.SUM)
        ms_e2e, ok_e2e = float(temax[0]), float(tesum[1])
    e2e = {"value": ok_e2e / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int((st_h.nbytes + xr_h.nbytes + sc_h.nbytes) * world),
           "d2h_bytes_per_step": int((th_h[0].nbytes + stt_h[0].nbytes) * world), "steps": e2e_steps,
           "api": "SpiralingController.step(state, ref, scenario=..., slot=i) -> thrust (pinned host tensors in / out), step i on stream i % S"}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    clocks_summary = clk.summary()
    props = torch.cuda.get_device_properties(devs)
    # ---- roofline of the dominant kernel, k_solve (rank 0's shard, last timed step)
    fl = algorithmic_flops(N, iters[:, 0], iters[:, 1])
    peak = C.c_double()
    L.check(eng.lib.ftmpc_fp64_peak(eng.handle, C.byref(peak), C.c_void_p(stream.cuda_stream)))
    solve_flops = fl["total"] - 5000.0 * B                      # everything but the allocation QP runs in k_solve
    achieved = solve_flops / (kernel_ms["k_solve"] * 1e-3) * 1e-12
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = algorithmic_bytes(N) * B / (ms_total / a.steps * 1e-3) * 1e-9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            # DRAM bytes of the ncu capture are per SOLVE (the capture runs a smaller batch): scale to this launch
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_solve") * B
        except Exception:
            traffic = None
    roofline = {"bound": "fp64", "kernel": "k_solve", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved / peak.value if peak.value > 0 else None, "traffic": traffic,
                "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, per solve, x instances of this launch",
                "peak_source": "FP64 FMA probe run live on this device (ftmpc_fp64_peak); MEASURED_PEAKS.json has no fp64 figure",
                "flops_model": "SURVEY.md 8d: K_sqp(F_lin+F_cond+F_chol)+K_qp F_iter+F_alloc with the kernel's own iteration counters",
                "algorithmic_flops_per_launch": solve_flops, "kernel_ms": kernel_ms,
                # share of k_solve in the device time of the step's own kernels (solo profiled step; with pipelined steps the
                # wall time per step is shorter than one launch, so the share is taken over the kernels, as the ncu list does)
                "kernel_share_of_step": kernel_ms["k_solve"] / (kernel_ms["k_solve"] + kernel_ms["k_alloc"]),
                "achieved_pipelined": solve_flops / (ms_total / a.steps * 1e-3) * 1e-12,
                "kernel_ms_note": "one step alone on the main stream after the timed region, in-kernel phase profile on",
                # cycles the persistent CTAs spent working (in-kernel clock64 marks, summed over CTAs) / (SMs x kernel time):
                # below 1 = SMs idling at the end of the launch while the last, hardest instances finish
                "sm_busy_frac": (float(cyc[:L.N_CYCLE_PHASES].sum()) / (props.multi_processor_count * kernel_ms["k_solve"] * 1e-3 * clocks_summary.get("sm_mhz", 1965.0) * 1e6)
                                 if cyc.sum() > 0 and clocks_summary.get("sm_mhz") else None),
                "phase_share": {n: round(float(c / cyc[:L.N_CYCLE_PHASES].sum()), 4) for n, c in zip(L.PHASE_NAMES[:L.N_CYCLE_PHASES], cyc)} if cyc.sum() > 0 else None,
                "phase_counters": {n: int(c) for n, c in zip(L.PHASE_NAMES[L.N_CYCLE_PHASES:], cyc[L.N_CYCLE_PHASES:])},
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                        "algorithmic_bytes_per_solve": algorithmic_bytes(N),
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"}}

    # ---- BASELINE configs[2] size as an extra key: 1024 instances per step.  A launch of 1024 is 7 instances per SM and ends on
    # its slowest instance (sm_busy_frac 0.4), so small batches are kept 4 deep in flight (independent batches, own buffers)
    small = None
    if hi - lo >= 1024 and world == 1 and not a.no_latency:
        Bs, Ss, Ks = 1024, 4, 16
        side4 = [torch.cuda.Stream(device=devs) for _ in range(Ss)]
        st_s, xr_s = st_d[:Bs].contiguous(), xr_d[:Bs].contiguous()
        sc_s = tuple(t[:Bs].contiguous() for t in sc_t)
        for i in range(2 * Ss):
            with torch.cuda.stream(side4[i % Ss]):
                eng.step(st_s, xr_s, scenario=sc_s, out=eng.buffers(Bs, 8 + i % Ss))
        torch.cuda.synchronize()
        ok_s = torch.zeros(Ss, dtype=torch.int64, device=devs)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for s_ in side4:
            s_.wait_event(s0)
        for i in range(Ks):
            with torch.cuda.stream(side4[i % Ss]):
                o_ = eng.step(st_s, xr_s, scenario=sc_s, out=eng.buffers(Bs, 8 + i % Ss))
                ok_s[i % Ss] += (o_["status"] == 0).sum()
        for s_ in side4:
            ev = torch.cuda.Event(); ev.record(s_); stream.wait_event(ev)
        s1.record(stream)
        torch.cuda.synchronize()
        small = {"value": float(ok_s.sum().item()) / (s0.elapsed_time(s1) * 1e-3), "unit": UNIT, "batch": Bs, "steps": Ks,
                 "streams": Ss, "ms_per_step": s0.elapsed_time(s1) / Ks,
                 "workload": "BASELINE configs[2] size: the first 1024 instances of the batch, device-resident, 4 batches in flight"}

    # ---- single-instance latency (BASELINE configs[1]): examples/sim.py default scenario, closed loop, warm start
    latency = None
    if not a.no_latency:
        from ft_mpc_b200.models import SpiralModel
        from ft_mpc_b200.util import BrokenThruster
        m1 = SystemModel(0.1)
        m1.set_fault(BrokenThruster(10, 1.0)); m1.set_fault(BrokenThruster(11, 1.0))
        c1 = SpiralingController(SpiralModel.from_system_model(m1), {"horizon": 15}, None, device=devs)
        c1.load_trajectory("hover", 30)
        from scipy.spatial.transform import Rotation
        x = np.concatenate([[1, 0, 1], [1, .5, 0], Rotation.from_euler("zyx", [50, 30, -10], degrees=True).as_quat(), [.3, .8, -.1]])
        lat, kms_l, its = [], [], []
        kbuf = (C.c_double * 2)()
        for k in range(300):                                       # the whole demo horizon (reactive.yaml:2,5; sim_env.py:109)
            t0 = time.perf_counter()
            u = c1.get_control(x, 0.1 * k + 1e-9)
            lat.append((time.perf_counter() - t0) * 1e3)
            its.append(c1.last_iters[0])
            x = m1.normalize_quaternion(m1.dynamics(x, u))
        # device time of k_solve alone for the same loop (events inside ftmpc_step; separate pass so that the event
        # bookkeeping does not sit in the latency numbers above)
        L.check(c1.engine.lib.ftmpc_profile_enable(c1.engine.handle, 1))
        for k in range(300, 340):
            u = c1.get_control(x, 0.1 * (k % 250) + 1e-9)
            L.check(c1.engine.lib.ftmpc_profile_read(c1.engine.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream), kbuf, None, 0))
            kms_l.append(float(kbuf[0]) + float(kbuf[1]))
            x = m1.normalize_quaternion(m1.dynamics(x, u))
        L.check(c1.engine.lib.ftmpc_profile_enable(c1.engine.handle, 0))
        cold, lat = lat[0], np.array(lat[1:])
        latency = {"p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)), "cold_first_ms": float(cold),
                   "steps": len(lat), "sqp_iters_mean": float(np.mean(its[1:])), "kernel_p50_ms": float(np.percentile(kms_l, 50)),
                   "workload": "examples/sim.py default scenario, N=15, closed loop over the 300-step demo horizon, warm start, "
                               "get_control(x,t) host call (pinned H2D of the state, one pinned D2H of the result record)"}

    # ---- CPU baseline (same algorithm on the host cores), bounded sample
    cpu = None
    if not a.no_cpu_baseline and world == 1:
        probe = min(B, 256)
        r0, *_ = cpu_solve_rate(N, cells, states, scen, xref, probe, 1, 0)
        sample = a.cpu_sample or int(min(B, max(probe, r0 * 12)))
        rate, ms, cores, okf = cpu_solve_rate(N, cells, states, scen, xref, sample, 1, 0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {sample} instances of the same batch, one pass, OpenMP over instances ({cores} threads), "
                         "C++ port of the same SQP on the reference NLP (oracle/cpu_port); the reference's casadi/IPOPT stack is not installable offline",
               "converged_frac": okf, **cpu_info()}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "converged_frac": ok_total / (Btot * a.steps),
            "sqp_iters_mean": float(iters[:, 0].mean()), "qp_iters_mean": float(iters[:, 1].mean()),
            "status_hist": np.bincount(status, minlength=6).tolist(), "clocks": clocks_summary, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "latency": latency, "batch_1024": small,
            "target": {"solves_per_s_8gpu": 1e6, "per_gpu": 125000.0}}
    _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _emit(line: dict) -> None:
    """the ONE JSON line goes to the process's original stdout; everything else written to fd 1 while the bench runs
    (NCCL prints its version banner there) has been sent to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1
if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sys.exit(main())
