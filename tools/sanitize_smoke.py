"""Tiny run of every kernel, meant to be executed under compute-sanitizer (memcheck / racecheck / synccheck)."""
import numpy as np
import torch

import bhmm_b200.hidden as hidden
from bhmm_b200 import _lib
from bhmm_b200.engine import TrajectoryBatch, unpack_stats
from bhmm_b200.util import testsystems as ts

import os
for N, T in [(3, 700), (10, 900), (16, 300), (40, 300)]:
    pi, A, means, sigmas, O, S = ts.gaussian_observations(N, 3, T, seed=N)
    from bhmm_b200.output_models import GaussianOutputModel
    pobs = GaussianOutputModel(N, means=means, sigmas=sigmas).p_obs(O[0])
    _lib.lib.bhmm_b200_set_chunking(128, 64)
    lp, alpha = hidden.forward(A, pobs, pi)
    beta = hidden.backward(A, pobs)
    gamma = hidden.state_probabilities(alpha, beta)
    hidden.state_counts(gamma, T)
    hidden.transition_counts(alpha, beta, A, pobs)
    hidden.viterbi(A, pobs, pi)
    hidden.sample_path(alpha, A, pobs, seed=1)
    _lib.lib.bhmm_b200_set_chunking(0, 0)
    b = TrajectoryBatch([O[0], O[1][:T // 2], O[2]], N, chunk=100, warm=50)
    st = unpack_stats(b.estep_gaussian(A, pi, means, sigmas).cpu().numpy(), N)
    b.viterbi_gaussian(A, pi, means, sigmas)
    b.gibbs_gaussian(A, pi, means, sigmas, seed=1, sweep=0)
    print(N, T, lp, st['loglik'], b.info())
    b.close()
pi, A, B, O, S = ts.discrete_observations(6, 30, 2, 500, seed=1)
b = TrajectoryBatch([O[0], O[1]], 6, chunk=100, warm=50)
stats, Bnum = b.estep_discrete(A, pi, B)
b.viterbi_discrete(A, pi, B)
b.gibbs_discrete(A, pi, B, seed=2, sweep=1)
print('discrete', float(stats[0]), float(Bnum.sum()))
torch.cuda.synchronize()
print('sanitize smoke done')
