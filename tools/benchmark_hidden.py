"""Per-function timing of the literal bhmm.hidden API, mirroring bhmm/tests/benchmark_hidden.py:62-72,163-196
(N=3, T=1e6 frames, per-function milliseconds), CUDA drop-in vs the reference C implementation on this host."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bhmm_b200.hidden as cuda
from bhmm_b200.output_models import GaussianOutputModel
from oracle.oracle import Oracle, have_reference_lib, build
build(ref=os.path.isdir('/root/reference'))
ref = Oracle('reference' if have_reference_lib() else 'port')

def best(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return 1e3 * min(ts)

for N, T in ((3, 1000000), (10, 100000)):
    rng = np.random.default_rng(0)
    X = rng.random((N, N)) + np.eye(N) * 10
    A = X / X.sum(axis=1)[:, None]
    pi = np.ones(N) / N
    means, sigmas = np.linspace(-1, 1, N), np.full(N, 0.5)
    obs = rng.integers(0, N, size=T).astype(float) * (2.0 / max(N - 1, 1)) - 1.0 + 0.5 * rng.standard_normal(T)
    gom = GaussianOutputModel(N, means=means, sigmas=sigmas)
    pobs = gom.p_obs(obs)
    lp, alpha = cuda.forward(A, pobs, pi)
    beta = cuda.backward(A, pobs)
    gamma = cuda.state_probabilities(alpha, beta)
    ao, bo, go, Co = np.zeros_like(alpha), np.zeros_like(alpha), np.zeros_like(alpha), np.zeros((N, N))
    rows = [
        ('p_obs', lambda: gom.p_obs(obs, out=go), lambda: ref.gaussian_p_obs(obs, means, sigmas)),
        ('forward', lambda: cuda.forward(A, pobs, pi, alpha_out=ao), lambda: ref.forward(A, pobs, pi)),
        ('backward', lambda: cuda.backward(A, pobs, beta_out=bo), lambda: ref.backward(A, pobs)),
        ('state_probabilities', lambda: cuda.state_probabilities(alpha, beta, gamma_out=go), lambda: ref.state_probabilities(alpha, beta)),
        ('state_counts', lambda: cuda.state_counts(gamma, T), lambda: ref.state_counts(gamma)),
        ('transition_counts', lambda: cuda.transition_counts(alpha, beta, A, pobs, out=Co), lambda: ref.transition_counts(alpha, beta, A, pobs)),
        ('viterbi', lambda: cuda.viterbi(A, pobs, pi), lambda: ref.viterbi(A, pobs, pi)),
        ('sample_path', lambda: cuda.sample_path(alpha, A, pobs, seed=1), lambda: ref.sample_path(alpha, A, seed=1)),
    ]
    print('N=%d T=%d   %-20s %10s %10s' % (N, T, 'function', 'cuda ms', 'ref C ms'))
    for name, f_cuda, f_ref in rows:
        print('            %-20s %10.2f %10.2f' % (name, best(f_cuda, 3), best(f_ref, 2)), flush=True)
