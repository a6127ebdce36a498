"""Tiny run of the panel / wide / register-Viterbi / chunked-Viterbi kernels, meant for compute-sanitizer
(memcheck, racecheck, synccheck):   BHMM_B200_PANEL=1 compute-sanitizer --tool racecheck python tools/sanitize_panel.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('BHMM_B200_PANEL', '1')
import numpy as np      # noqa: E402
import torch            # noqa: E402

from bhmm_b200.engine import TrajectoryBatch, unpack_stats   # noqa: E402

rng = np.random.default_rng(7)
for N in (32, 100, 37, 64, 21):
    X = rng.random((N, N)) ** 2 + 1e-3
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-5, 5, N), np.linspace(0.5, 2.0, N)
    obs = []
    for Tk in (160, 97, 1, 230, 64, 33, 120, 75, 140, 50):
        s = rng.integers(0, N, size=Tk)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(Tk))
    b = TrajectoryBatch(obs, N, chunk=60, warm=30)
    st = unpack_stats(b.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    b.viterbi_gaussian(A, pi, means, sigmas)
    print(N, 'gaussian', st['loglik'], b.info(), flush=True)
    b.close()
    M = 40
    B = rng.random((N, M)) ** 3 + 1e-4
    B /= B.sum(axis=1)[:, None]
    sym = [rng.integers(0, M, size=Tk).astype(np.int32) for Tk in (150, 80, 17)]
    b = TrajectoryBatch(sym, N, chunk=50, warm=20)
    stats, Bnum = b.estep_discrete(A, pi, B)
    b.viterbi_discrete(A, pi, B)
    print(N, 'discrete', float(stats[0]), float(Bnum.sum()), flush=True)
    b.close()
torch.cuda.synchronize()
print('sanitize panel done')
