"""Probe: why is the device-resident loop of bench.py slower than the e2e loop?  Variants of the timed loop on one GPU."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
import ftmpc_import; ftmpc_import.load()
from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
from ft_mpc_b200.models import SystemModel
N, B, K, S = 20, 8192, 12, 4
cells, states, scen, xref = bench.make_workload(B, N, 1)
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, cells, device=dev)
f64 = torch.float64
st_d = torch.tensor(states, dtype=f64, device=dev); xr_d = torch.tensor(xref, dtype=f64, device=dev)
sc_idx = torch.tensor(scen, device=dev); sc_t = eng.scenario_tensors(sc_idx)
stream = torch.cuda.current_stream(); side = [torch.cuda.Stream(device=dev) for _ in range(S)]
for i in range(2 * S):
    with torch.cuda.stream(side[i % S]): eng.step(st_d, xr_d, scenario=sc_t, out=eng.buffers(B, i % S))
torch.cuda.synchronize()
ok_acc = torch.zeros(S, dtype=torch.int64, device=dev)
def run(variant):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for s_ in side: s_.wait_event(e0)
    for i in range(K):
        with torch.cuda.stream(side[i % S]):
            if variant == "tuple": eng.step(st_d, xr_d, scenario=sc_t, out=eng.buffers(B, i % S))
            elif variant == "index": eng.step(st_d, xr_d, scenario=sc_idx, out=eng.buffers(B, i % S))
            elif variant == "fresh_inputs": eng.step(st_d.clone(), xr_d.clone(), scenario=sc_t, out=eng.buffers(B, i % S))
            elif variant == "sleep": eng.step(st_d, xr_d, scenario=sc_t, out=eng.buffers(B, i % S)); time.sleep(0.05)
            elif variant == "okacc":
                o = eng.step(st_d, xr_d, scenario=sc_t, out=eng.buffers(B, i % S)); ok_acc[i % S] += (o["status"] == 0).sum()
            elif variant == "okacc_launches":
                o = eng.step(st_d, xr_d, scenario=sc_t, out=eng.buffers(B, i % S)); ok_acc[i % S] += (o["status"] == 0).sum()
                eng.lib.ftmpc_last_launches(eng.handle)
    for s_ in side:
        ev = torch.cuda.Event(); ev.record(s_); stream.wait_event(ev)
    e1.record(stream); torch.cuda.synchronize()
    return B * K / (e0.elapsed_time(e1) * 1e-3)
for rep in range(2):
    for v in ("tuple", "okacc", "okacc_launches"):
        print(v, round(run(v)))
