"""The REFERENCE's MaximumLikelihoodEstimator (unchanged, imported from baseline/_ref) on 'c', on 'cuda' through the literal
five-copy path, and on 'cuda' with the buffer-identity device cache: seconds per EM iteration.

    python tools/reference_cache_probe.py [--nstates 10] [--trajectories 16] [--frames 100000] [--iters 3]
"""
import argparse, json, os, sys, time, warnings
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import msmtools_stub
msmtools_stub.install()
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
warnings.simplefilter('ignore')
import bhmm
import bhmm_b200
from bhmm_b200.hidden import api as cuda_api
from bhmm_b200.util import testsystems as ts
from bhmm.util import config
from bhmm.hmm.generic_hmm import HMM
from bhmm.output_models.gaussian import GaussianOutputModel
from bhmm.estimators.maximum_likelihood import MaximumLikelihoodEstimator

ap = argparse.ArgumentParser()
ap.add_argument('--nstates', type=int, default=10)
ap.add_argument('--trajectories', type=int, default=16)
ap.add_argument('--frames', type=int, default=100000)
ap.add_argument('--iters', type=int, default=3)
args = ap.parse_args()
N, K, T = args.nstates, args.trajectories, args.frames
pi, A, means, sigmas, O, S = ts.gaussian_observations(N, K, T, seed=3)
pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
obs = [O[k] for k in range(K)]


def fit(kernel, cache):
    bhmm_b200.install(bhmm, device_cache=cache)
    cuda_api.set_device_cache(cache)
    config.kernel = kernel
    init = HMM(pi0, A0, GaussianOutputModel(N, means=m0.copy(), sigmas=s0.copy()))
    est = MaximumLikelihoodEstimator(obs, N, initial_model=init, reversible=False, stationary=False, accuracy=-np.inf, maxit=args.iters)
    t0 = time.perf_counter()
    est.fit()
    dt = time.perf_counter() - t0
    return dt / args.iters, est.likelihoods[-1]

fit('cuda', True)                                   # warm-up (CUDA context, allocator)
out = {}
for name, kernel, cache in (('c', 'c', False), ('cuda_literal', 'cuda', False), ('cuda_cached', 'cuda', True)):
    s, ll = fit(kernel, cache)
    out[name] = {'seconds_per_iteration': s, 'frames_iters_per_s': K * T / s, 'loglik': float(ll)}
out['cache_stats'] = cuda_api.device_cache_stats()
out['workload'] = '%d states, %d trajectories x %d frames, reference MaximumLikelihoodEstimator.fit (incl. its final Viterbi)' % (N, K, T)
print(json.dumps(out))
