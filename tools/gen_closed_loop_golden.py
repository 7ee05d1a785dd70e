#!/usr/bin/env python
"""Generate tests/golden/closed_loop_N15.npz: the ORACLE's closed loop of the examples/sim.py default scenario
(faults 10, 11 stuck fully open, hover at the origin, N = 15; SimulationEnvironment.step, sim_env.py:77-99) for
STEPS steps, once with the noise switched off and once with the recorded noise tensor SURVEY.md 8(d) config 1 names
(numpy.random.default_rng(0).uniform(0, 1e-3, (300, 13)), added to all 13 states after the plant step, then the
quaternion is re-normalised: sim_env.py:88-93).  Every step: oracle.solve_nlp on the pinned NLP (warm start = the
previous solution shifted by one stage with a zero tail, spiraling_mpc.py:324-331), oracle.get_control, oracle plant.
Usage: python tools/gen_closed_loop_golden.py [steps]        (two processes, ~10-20 minutes)
"""
import sys
import time
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import ftmpc_oracle as o  # noqa: E402

N = 15
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def noise_tensor():
    return np.random.default_rng(0).uniform(0.0, 1e-3, (300, 13))


def run(noisy):
    prob, x = o.default_problem(N)
    fs = prob.fs
    traj, nom = o.assign_trajectory(o.hover_trajectory(30, 0.1), N, 0.1)
    noise = noise_tensor()
    xs, Us, ths, fsum, kkts = [], [], [], [], []
    Uprev = None
    for step in range(STEPS):
        t0 = time.time()
        xr, ur = o.window(traj, nom, step, N)
        p = o.Problem(fs, N, o.robot_to_center(x), xr, ur)
        U0 = None if Uprev is None else np.vstack([Uprev[1:], np.zeros((1, 6))])
        res = o.get_control(p, o.solve_nlp(p, U0=U0))
        assert res["kkt_stat"] < 1e-7 and res["kkt_viol"] < 1e-8 and res["alloc_ok"], (step, res["kkt_stat"], res["kkt_viol"])
        xs.append(x.copy()); Us.append(res["U"].copy()); ths.append(res["thrust"].copy()); fsum.append(res["f"])
        kkts.append([res["kkt_stat"], res["kkt_viol"]])
        x = o.plant_rk4(x, res["thrust"], fs, 0.1)                              # sim_env.py:85
        if noisy:
            x = x + noise[step]                                                 # sim_env.py:88-91
        x = o.normalize_quaternion_robot(x)                                     # sim_env.py:93
        Uprev = res["U"]
        print(f"noisy={noisy} step {step}: f={res['f']:.6f} kkt={res['kkt_stat']:.1e} {time.time() - t0:.0f}s", flush=True)
    xs.append(x.copy())
    return dict(x=np.array(xs), U=np.array(Us), thrust=np.array(ths), f=np.array(fsum), kkt=np.array(kkts))


def main():
    with Pool(2) as p:
        clean, noisy = p.map(run, [False, True])
    out = {f"clean::{k}": v for k, v in clean.items()}
    out.update({f"noisy::{k}": v for k, v in noisy.items()})
    out["N"] = np.array(N)
    out["noise_seed"] = np.array(0)
    dst = ROOT / "tests" / "golden" / "closed_loop_N15.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst)


if __name__ == "__main__":
    main()
