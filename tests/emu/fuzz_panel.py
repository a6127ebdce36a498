"""Randomised runs of the emulated panel-family kernels against the plain forward-backward (not collected by pytest):

    python tests/emu/fuzz_panel.py [seconds] [seed]

Random state counts (17..104, and 32 on both kernel variants), ragged trajectories (1..90 frames), random chunk lengths,
grids and emission kinds.  Prints one line per case; exits non-zero at the first mismatch."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ctypes as C                                     # noqa: E402
import test_panel_emulated_cpu as T                   # noqa: E402

emu = C.CDLL(os.path.join(HERE, 'panel_emu.so'))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
case = 0
while time.time() < t_end:
    case += 1
    N = int(rng.choice([32, 32, int(rng.integers(17, 105))]))
    wide32 = bool(N == 32 and rng.random() < 0.5)
    emu.panel_emu_force_wide(1 if (wide32 or N != 32) else 0)
    r2, A, pi, mu, sigma = T.model(int(rng.integers(1 << 30)), mixing=float(rng.uniform(2.0, 5.0)), N=N)
    K = int(rng.integers(1, 7))
    Ts = [int(rng.integers(1, 91)) for _ in range(K)]
    chunk = int(rng.integers(5, 41))
    warm = int(rng.integers(45, 70))
    grid = int(rng.integers(1, 4))
    plan = T.make_plan(Ts, chunk)
    kind = int(rng.choice([T.EM_GAUSS, T.EM_DISC]))
    if kind == T.EM_GAUSS:
        trajs = [mu[r2.integers(0, N, t)] + 0.7 * r2.standard_normal(t) for t in Ts]
        obs = np.concatenate(trajs)
        ps = [T.gauss(o[:, None], mu[None, :], sigma[None, :]) for o in trajs]
        run = T.Run(emu, plan, A, pi, kind, obs=obs, mu=mu, sigma=sigma, ignore_outliers=1)
    else:
        M = int(rng.integers(2, 20))
        B = r2.random((N, M)) ** 2 + 1e-3
        B /= B.sum(axis=1)[:, None]
        syms = [r2.integers(0, M, t).astype(np.int32) for t in Ts]
        ps = [B[:, s].T.copy() for s in syms]
        run = T.Run(emu, plan, A, pi, kind, sym=np.concatenate(syms), Bt=np.ascontiguousarray(B.T), M=M)
    run.wide32 = wide32
    ref = T.plain_estep(ps, A, pi)
    run.forward(grid, warm=warm)
    st = run.backward_stats(grid, warm, run.alpha)
    ea = float(np.max(np.abs(run.alpha - ref['alpha'])))
    ec = float(np.max(np.abs(st['C'] - ref['C']) / (ref['C'] + 1e-300)))
    el = abs(run.chain_ll.sum() - ref['ll']) / abs(ref['ll'])
    ok = ea < 1e-9 and ec < 1e-7 and el < 1e-10 and abs(st['C'].sum() - (sum(Ts) - K)) < 1e-7
    print('case %d: N=%d%s kind=%d K=%d chunk=%d warm=%d grid=%d chains=%d: alpha %.1e C %.1e ll %.1e %s'
          % (case, N, ' (wide)' if wide32 else '', kind, K, chunk, warm, grid, len(plan[0]), ea, ec, el, 'ok' if ok else 'FAIL'),
          flush=True)
    if not ok:
        sys.exit(1)
print('%d cases ok' % case)
