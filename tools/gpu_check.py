#!/usr/bin/env python
"""GPU bring-up check (development aid): GPU path vs the CPU port of the same algorithm on a scenario batch,
plus CUDA-event timing of ftmpc_step.  Usage: python tools/gpu_check.py [B] [N]"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ftmpc_import  # noqa: E402

ftmpc_import.load()
from ft_mpc_b200 import _lib as L  # noqa: E402
from ft_mpc_b200.controllers.spiraling_mpc import BatchedMPC, DEFAULT_Q, DEFAULT_R  # noqa: E402
from ft_mpc_b200.models import SystemModel  # noqa: E402
from ft_mpc_b200.util import scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cells = scenarios.load_cells(kinds=("single", "double"))
model = SystemModel(0.1)
eng = BatchedMPC(model, N, DEFAULT_Q, DEFAULT_R, cells)
states = scenarios.random_states(B, 1)
scen = np.arange(B) % len(cells)
xref = scenarios.hover_reference(B, N)
dev = eng.device
st_d = torch.tensor(states, device=dev)
xr_d = torch.tensor(xref, device=dev)
sc_d = torch.tensor(scen, device=dev)
sc_t = eng.scenario_tensors(sc_d)
out = eng.step(st_d, xr_d, scenario=sc_t)
torch.cuda.synchronize()
status = out["status"].cpu().numpy()
iters = out["iters"].cpu().numpy()
print("GPU status hist", np.bincount(status, minlength=5), "sqp it mean/max", iters[:, 0].mean(), iters[:, 0].max(),
      "qp it mean", iters[:, 1].mean())
# timing
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(st_d, xr_d, scenario=sc_t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"B={B} N={N}: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} solves/s")
# CPU port on a subset
cpu = C.CDLL(str(ROOT / "oracle" / "_cpu" / "libftmpc_cpu.so"))
nb = min(B, 64)
p = lambda a: a.ctypes.data_as(C.c_void_p)
mask = np.array([eng.mask_tab[s] for s in scen[:nb]], np.uint16)
ff = np.ascontiguousarray(np.stack([eng.fault_force_tab[s] for s in scen[:nb]]))
hidx = scen[:nb].astype(np.int32)
zw = np.zeros((nb, eng.nz)); th = np.zeros((nb, 16)); u0 = np.zeros((nb, 6))
act = np.zeros((nb, (eng.mc + 31) // 32), np.uint32); stc = np.zeros(nb, np.int32); itc = np.zeros((nb, 2), np.int32)
cost = np.zeros(nb)
t = time.time()
cpu.ftmpc_cpu_step(C.byref(eng.cfg), p(eng.hull_table), nb, p(np.ascontiguousarray(states[:nb])),
                   p(np.ascontiguousarray(xref[:nb])), None, p(mask), p(ff), p(hidx), 0, p(zw), p(th), p(u0), p(act),
                   p(stc), p(itc), p(cost), None, 0)
print(f"CPU port: {nb} solves in {time.time() - t:.2f} s")
g_u0 = out["u0"][:nb].cpu().numpy(); g_th = out["thrust"][:nb].cpu().numpy()
g_act = out["active"][:nb].cpu().numpy().view(np.uint32)
okm = (status[:nb] == 0) & (stc == 0)
print("status equal", (status[:nb] == stc).all(), "both ok", okm.sum())
print("max |u0 gpu-cpu|", np.abs(g_u0 - u0)[okm].max(), "max |thrust gpu-cpu|", np.abs(g_th - th)[okm].max(),
      "active sets equal", (g_act[okm] == act[okm]).all(), "iters equal", (iters[:nb] == itc).all())
