"""debug aid: run golden cases one by one on the GPU and print the solver scalars left in the CTA slot"""
import sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import ftmpc_import; ftmpc_import.load()
import helpers as H
from ft_mpc_b200.controllers.spiraling_mpc import DEFAULT_Q, DEFAULT_R, BatchedMPC
from ft_mpc_b200.models import SystemModel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 15
method = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = np.load('tests/golden/nlp_cases.npz')
names = ['F', 'CSUM', 'NU', 'GD', 'THETA', 'DMAX', 'DELTA', 'LAMMAX', 'STATUS', 'ITER', 'QPIT', 'NACT', 'CHOLFAIL', 'QPST', 'ALPHA', 'CMAX', 'SIGMA', 'HFAIL', 'DPREV', 'DREF', 'FREF']
n, nv, mc = 6 * N, 6 * N + 1, 26 * N + 72
m = mc + 2
oSc = n + nv + (N + 1) * 13 + mc + 2 * m + (N + 1) * 13 + 169 * N * 2 + 9 + 81
for k in H.cases_with_horizon(g, N):
    if g['warm'][k]: continue
    fs = H.case_faults(g, k)
    for it in ([1, 2, 3, 60] if len(sys.argv) > 3 else [60]):
        eng = BatchedMPC(SystemModel(0.1), N, DEFAULT_Q, DEFAULT_R, [fs], qp_method=method, max_sqp_iter=it)
        d = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')
        out = eng.step(d(g['x0'][[k]]), d(g['xref'][[k], :N + 1]))
        torch.cuda.synchronize()
        sc = out['ws'].view(torch.float64)[oSc:oSc + len(names)].cpu().numpy()
        print(g['name'][k], 'maxit', it, 'status', int(out['status'][0]), 'iters', out['iters'][0].tolist(),
              {nm: float('%.6g' % v) for nm, v in zip(names, sc) if nm in ('F', 'CSUM', 'THETA', 'DMAX', 'DELTA', 'QPIT', 'NACT', 'CHOLFAIL', 'QPST', 'ALPHA', 'SIGMA')}, flush=True)
