import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['BHMM_B200_FAMILY'] = sys.argv[1] if len(sys.argv) > 1 else 'team'
import numpy as np, torch
from bhmm_b200 import _lib
from bhmm_b200.engine import TrajectoryBatch
N = 4
rng = np.random.default_rng(904)
eps = 2e-5
A = (1.0 - eps) * np.eye(N) + eps * np.ones((N, N)) / N
A /= A.sum(axis=1)[:, None]
pi = rng.random(N) + 0.5; pi /= pi.sum()
means, sigmas = np.linspace(-0.02, 0.02, N), np.ones(N)
obs = [rng.standard_normal(T) for T in (9000, 4000, 300)]
b = TrajectoryBatch(obs, N, chunk=400, warm=32)
try:
    b.estep_gaussian(A, pi, means, sigmas)
    print('ok', b.info(), b.exact_scans)
except Exception as e:
    print('ERR', str(e)[:150])
n = b.info()['chains']
for d in (1, -1):
    u = np.zeros((n, N)); e = np.zeros((n, N))
    _lib.lib.bhmm_b200_batch_debug_handovers(b._handle, d, _lib.dptr(u), _lib.dptr(e))
    for c in range(n):
        nb = c - 1 if d > 0 else c + 1
        if 0 <= nb < n:
            mm = np.max(np.abs(u[c] - e[nb]) / np.maximum(np.abs(u[c]), 1e-300))
            print(d, c, 'used', np.round(u[c], 4), 'nb end', np.round(e[nb], 4), 'mismatch %.2e' % mm)
