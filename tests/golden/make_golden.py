"""Generate tests/golden/*.npz from the REFERENCE itself (run in the build container only).

TEST INFRASTRUCTURE.  The reference package is imported from a scratch copy whose three hot-path
Cython extensions were built in place (bhmm.hidden.impl_c.hidden, bhmm.output_models.impl_c.
gaussian / discrete):

    cp -r /root/reference /tmp/refbuild && cd /tmp/refbuild && python build_min.py   
    python tests/golden/make_golden.py /tmp/refbuild

`msmtools` (absent, un-pinned) is replaced by tests/golden/msmtools_stub.py, which only supports the
non-reversible estimators; every model below therefore uses an asymmetric transition matrix so
the reference's EM takes `estimate_P(..., reversible=False)` (maximum_likelihood.py:307).

All outputs come from the reference's public API with config.kernel='c':
bhmm.hidden.{forward,backward,state_probabilities,state_counts,transition_counts,viterbi,
sample_path}, OutputModel.p_obs, MaximumLikelihoodEstimator.fit, BayesianHMMSampler internals.
The fixtures travel to the GPU box; /root/reference does not.
"""
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import msmtools_stub  # noqa: E402

msmtools_stub.install()
REF = sys.argv[1] if len(sys.argv) > 1 else '/tmp/refbuild'
sys.path.insert(0, REF)
warnings.simplefilter('ignore')

import bhmm  # noqa: E402
from bhmm.hidden import api as hidden  # noqa: E402
from bhmm.util import config  # noqa: E402
from bhmm.output_models.gaussian import GaussianOutputModel  # noqa: E402
from bhmm.output_models.discrete import DiscreteOutputModel  # noqa: E402
from bhmm.hmm.generic_hmm import HMM  # noqa: E402
from bhmm.estimators.maximum_likelihood import MaximumLikelihoodEstimator  # noqa: E402

config.kernel = 'c'
hidden.set_implementation('c')


def transition_matrix(n, rng, lmin=10.0, lmax=100.0, symmetric=False):
    """Vectorised restatement of the recipe of bhmm/util/testsystems.py:26-65 (seeded)."""
    lt = np.linspace(math.log(lmin), math.log(lmax), n)
    diag = 1.0 - 1.0 / np.exp(lt)
    X = rng.random((n, n))
    if symmetric:
        X = X + X.T
    T = X / X.sum(axis=1)[:, None]
    for i in range(n):
        T[i, i] = 0
        T[i, :] *= (1.0 - diag[i]) / T[i, :].sum()
        T[i, i] = 1.0 - T[i, :].sum()
    return T


def simulate(A, pi, T, rng):
    n = len(pi)
    s = np.zeros(T, dtype=np.int64)
    s[0] = rng.choice(n, p=pi)
    cum = np.cumsum(A, axis=1)
    u = rng.random(T)
    for t in range(1, T):
        s[t] = min(np.searchsorted(cum[s[t - 1]], u[t]), n - 1)
    return s


def hidden_suite(A, pi, pobs, seed):
    """Every bhmm.hidden function on one (A, pi, pobs), reference impl 'c'."""
    logprob, alpha = hidden.forward(A, pobs, pi)
    beta = hidden.backward(A, pobs)
    gamma = hidden.state_probabilities(alpha, beta)
    counts = hidden.state_counts(gamma, pobs.shape[0])
    Cm = hidden.transition_counts(alpha, beta, A, pobs)
    vpath = hidden.viterbi(A, pobs, pi)
    spath = hidden.sample_path(alpha, A, pobs, seed=seed)
    return dict(A=A, pi=pi, pobs=pobs, logprob=np.float64(logprob), alpha=alpha, beta=beta, gamma=gamma,
                counts=counts, C=Cm, viterbi=vpath, sample_seed=np.int64(seed), sample_path=spath)


def main():
    out = {}

    # 1. toy model, bhmm/tests/test_hidden.py:58-71
    A = np.array([[0.9, 0.1], [0.1, 0.9]])
    pi = np.array([0.5, 0.5])
    pobs = np.array([[0.1, 0.9]] * 4 + [[0.5, 0.5]] + [[0.9, 0.1]] * 5)
    out['hidden_toy'] = hidden_suite(A, pi, pobs, 42)

    # 2. three-state Gaussian example, bhmm/tests/test_hidden.py:74-84 (observations seeded here)
    rng = np.random.default_rng(11)
    A = np.array([[0.97, 0.02, 0.01], [0.1, 0.8, 0.1], [0.01, 0.02, 0.97]])
    pi = np.array([0.45, 0.1, 0.45])
    means, sigmas = np.array([-1.0, 0.0, 1.0]), np.array([0.5, 0.5, 0.5])
    gom = GaussianOutputModel(3, means=means, sigmas=sigmas)
    gom.set_implementation('c')
    obs = rng.integers(0, 3, size=2000).astype(np.float64)
    d = hidden_suite(A, pi, gom.p_obs(obs), 7)
    d.update(obs=obs, means=means, sigmas=sigmas)
    out['hidden_gauss3'] = d

    # 3. ten-state dalton-style Gaussian model (testsystems.py:105-188 recipe), incl. two far outliers
    rng = np.random.default_rng(3)
    n = 10
    A = transition_matrix(n, rng)
    pi = msmtools_stub.stationary_distribution(A)
    means, sigmas = np.linspace(-5, 5, n), np.linspace(0.5, 2.0, n)
    s = simulate(A, pi, 1500, rng)
    obs = means[s] + sigmas[s] * rng.standard_normal(1500)
    obs[700] = 400.0      # every state's density underflows to 0 -> outlier rule (outputmodel.py:119-131)
    obs[1203] = -80.0
    gom = GaussianOutputModel(n, means=means, sigmas=sigmas)
    gom.set_implementation('c')
    d = hidden_suite(A, pi, gom.p_obs(obs), 123)
    d.update(obs=obs, means=means, sigmas=sigmas, states=s)
    out['hidden_dalton10'] = d

    # 4. discrete output model, N=6 states, M=30 symbols
    rng = np.random.default_rng(4)
    n, m = 6, 30
    A = transition_matrix(n, rng)
    pi = msmtools_stub.stationary_distribution(A)
    centers = np.linspace(2, m - 3, n)
    B = np.exp(-0.5 * ((np.arange(m)[None, :] - centers[:, None]) / 2.5) ** 2) + 1e-4
    B /= B.sum(axis=1)[:, None]
    s = simulate(A, pi, 1200, rng)
    dobs = np.array([rng.choice(m, p=B[k]) for k in s], dtype=np.int32)
    dom = DiscreteOutputModel(B)
    dom.set_implementation('c')
    d = hidden_suite(A, pi, np.ascontiguousarray(dom.p_obs(dobs)), 5)
    d.update(obs=dobs, B=B, states=s)
    out['hidden_discrete'] = d

    # 5. Baum-Welch through the reference estimator: 3-state Gaussian, 4 ragged trajectories,
    #    non-reversible initial model, exactly 6 iterations (accuracy=-inf disables the stop test,
    #    maximum_likelihood.py:389-394).
    rng = np.random.default_rng(5)
    n = 3
    Atrue = transition_matrix(n, rng)
    pitrue = msmtools_stub.stationary_distribution(Atrue)
    mt, st = np.linspace(-5, 5, n), np.linspace(0.5, 2.0, n)
    lengths = [1500, 900, 1201, 333]
    observations = []
    for L in lengths:
        s = simulate(Atrue, pitrue, L, rng)
        observations.append(mt[s] + st[s] * rng.standard_normal(L))
    A0 = np.array([[0.90, 0.06, 0.04], [0.05, 0.90, 0.05], [0.03, 0.07, 0.90]])
    pi0 = np.ones(n) / n
    m0, s0 = mt + 0.5, np.ones(n)
    init = HMM(pi0, A0, GaussianOutputModel(n, means=m0.copy(), sigmas=s0.copy()))
    est = MaximumLikelihoodEstimator(observations, n, initial_model=init, reversible=False, stationary=False,
                                     accuracy=-np.inf, maxit=6)
    model = est.fit()
    d = dict(A0=A0, pi0=pi0, means0=m0, sigmas0=s0, lengths=np.array(lengths),
             likelihoods=np.array(est.likelihoods), A=model.transition_matrix, pi=model.initial_distribution,
             means=model.output_model.means, sigmas=model.output_model.sigmas,
             count_matrix=est.count_matrix, initial_count=est.initial_count)
    for k, o in enumerate(observations):
        d['obs%d' % k] = o
        d['viterbi%d' % k] = np.asarray(model.hidden_state_trajectories[k])
    out['em_gauss3'] = d

    # 6. Baum-Welch, discrete output model (N=4, M=12), 3 trajectories, 4 iterations
    rng = np.random.default_rng(6)
    n, m = 4, 12
    Atrue = transition_matrix(n, rng)
    pitrue = msmtools_stub.stationary_distribution(Atrue)
    centers = np.linspace(1, m - 2, n)
    Btrue = np.exp(-0.5 * ((np.arange(m)[None, :] - centers[:, None]) / 1.2) ** 2) + 1e-3
    Btrue /= Btrue.sum(axis=1)[:, None]
    lengths = [800, 1100, 257]
    dobservations = []
    for L in lengths:
        s = simulate(Atrue, pitrue, L, rng)
        dobservations.append(np.array([rng.choice(m, p=Btrue[k]) for k in s], dtype=np.int32))
    A0 = transition_matrix(n, np.random.default_rng(66), lmin=5, lmax=20)
    pi0 = np.ones(n) / n
    B0 = np.exp(-0.5 * ((np.arange(m)[None, :] - centers[:, None]) / 2.5) ** 2) + 1e-2
    B0 /= B0.sum(axis=1)[:, None]
    init = HMM(pi0, A0, DiscreteOutputModel(B0.copy()))
    est = MaximumLikelihoodEstimator(dobservations, n, initial_model=init, reversible=False, stationary=False,
                                     accuracy=-np.inf, maxit=4)
    model = est.fit()
    d = dict(A0=A0, pi0=pi0, B0=B0, lengths=np.array(lengths), likelihoods=np.array(est.likelihoods),
             A=model.transition_matrix, pi=model.initial_distribution,
             B=model.output_model.output_probabilities, count_matrix=est.count_matrix,
             initial_count=est.initial_count)
    for k, o in enumerate(dobservations):
        d['obs%d' % k] = o
        d['viterbi%d' % k] = np.asarray(model.hidden_state_trajectories[k])
    out['em_discrete'] = d

    # 7. One Gibbs hidden-path update in the reference's call sequence
    #    (bayesian_sampling.py:283-331: p_obs -> forward -> sample_path(seed)) plus the path
    #    statistics the parameter draws consume (generic_hmm.py:297-334,398-431).
    rng = np.random.default_rng(7)
    n = 3
    A = transition_matrix(n, rng)
    pi = msmtools_stub.stationary_distribution(A)
    means, sigmas = np.linspace(-5, 5, n), np.linspace(0.5, 2.0, n)
    gom = GaussianOutputModel(n, means=means, sigmas=sigmas)
    gom.set_implementation('c')
    lengths = [700, 450, 1000]
    d = dict(A=A, pi=pi, means=means, sigmas=sigmas, lengths=np.array(lengths), seed=np.int64(99))
    paths, obs_list = [], []
    for k, L in enumerate(lengths):
        s = simulate(A, pi, L, rng)
        o = means[s] + sigmas[s] * rng.standard_normal(L)
        pobs = gom.p_obs(o)
        alpha = hidden.forward(A, pobs, pi)[1]
        path = hidden.sample_path(alpha, A, pobs, seed=99)
        d['obs%d' % k], d['alpha%d' % k], d['path%d' % k] = o, alpha, path
        paths.append(path)
        obs_list.append(o)
    hm = HMM(pi, A, gom)
    hm.hidden_state_trajectories = paths
    d['count_matrix'] = hm.count_matrix()
    d['count_init'] = hm.count_init()
    for i in range(n):
        oi = hm.collect_observations_in_state(obs_list, i)
        d['obs_in_state_n%d' % i] = np.int64(len(oi))
        d['obs_in_state_mean%d' % i] = np.float64(np.mean(oi))
        d['obs_in_state_msd%d' % i] = np.float64(np.mean((oi - means[i]) ** 2))
    out['gibbs_gauss3'] = d

    # 8. The public entry point bhmm.estimate_hmm (api.py:309-372) with an initial model (the reference's own initial-model
    #    heuristics need msmtools): lagged observations (api.py:70-94), the estimator's real convergence test
    #    (accuracy=1e-3), non-reversible; Gaussian with lag 3 and discrete with lag 2.  Also lag_observations with a stride.
    rng = np.random.default_rng(8)
    n = 3
    Atrue = transition_matrix(n, rng, lmin=30.0, lmax=200.0)
    pitrue = msmtools_stub.stationary_distribution(Atrue)
    mt, st = np.array([-3.0, 0.5, 4.0]), np.array([0.8, 1.1, 0.6])
    lengths = [2400, 1501, 999, 3]
    observations = []
    for L in lengths:
        s = simulate(Atrue, pitrue, L, rng)
        observations.append(mt[s] + st[s] * rng.standard_normal(L))
    A0 = np.array([[0.80, 0.15, 0.05], [0.10, 0.80, 0.10], [0.04, 0.16, 0.80]])
    init = HMM(np.ones(n) / n, A0, GaussianOutputModel(n, means=mt - 0.4, sigmas=np.ones(n)))
    model = bhmm.estimate_hmm(observations, n, lag=3, initial_model=init, reversible=False, stationary=False,
                              accuracy=1e-3, maxit=200)
    lagged = bhmm.lag_observations(observations, 3)
    strided = bhmm.lag_observations(observations, 4, stride=2)
    d = dict(A0=A0, pi0=np.ones(n) / n, means0=mt - 0.4, sigmas0=np.ones(n), lengths=np.array(lengths), lag=np.int64(3),
             A=model.transition_matrix, pi=model.initial_distribution, means=model.output_model.means,
             sigmas=model.output_model.sigmas, likelihood=np.float64(model.likelihood), model_lag=np.int64(model.lag),
             n_lagged=np.int64(len(lagged)), lagged_lengths=np.array([len(o) for o in lagged]),
             strided_lengths=np.array([len(o) for o in strided]),
             strided_heads=np.array([o[0] for o in strided]), strided_tails=np.array([o[-1] for o in strided]))
    for k, o in enumerate(observations):
        d['obs%d' % k] = o
    for k, pth in enumerate(model.hidden_state_trajectories):
        d['viterbi%d' % k] = np.asarray(pth)
    out['api_estimate_gauss3'] = d

    rng = np.random.default_rng(9)
    n, m = 3, 9
    Atrue = transition_matrix(n, rng, lmin=20.0, lmax=100.0)
    pitrue = msmtools_stub.stationary_distribution(Atrue)
    centers = np.linspace(1, m - 2, n)
    Btrue = np.exp(-0.5 * ((np.arange(m)[None, :] - centers[:, None]) / 1.0) ** 2) + 1e-3
    Btrue /= Btrue.sum(axis=1)[:, None]
    lengths = [1800, 1203]
    dobservations = []
    for L in lengths:
        s = simulate(Atrue, pitrue, L, rng)
        dobservations.append(np.array([rng.choice(m, p=Btrue[k]) for k in s], dtype=np.int32))
    A0 = transition_matrix(n, np.random.default_rng(99), lmin=5, lmax=20)
    B0 = np.exp(-0.5 * ((np.arange(m)[None, :] - centers[:, None]) / 2.0) ** 2) + 1e-2
    B0 /= B0.sum(axis=1)[:, None]
    init = HMM(np.ones(n) / n, A0, DiscreteOutputModel(B0.copy()))
    model = bhmm.estimate_hmm(dobservations, n, lag=2, initial_model=init, reversible=False, stationary=False,
                              accuracy=1e-3, maxit=200)
    d = dict(A0=A0, pi0=np.ones(n) / n, B0=B0, lengths=np.array(lengths), lag=np.int64(2), A=model.transition_matrix,
             pi=model.initial_distribution, B=model.output_model.output_probabilities,
             likelihood=np.float64(model.likelihood), model_lag=np.int64(model.lag))
    for k, o in enumerate(dobservations):
        d['obs%d' % k] = o
    for k, pth in enumerate(model.hidden_state_trajectories):
        d['viterbi%d' % k] = np.asarray(pth)
    out['api_estimate_discrete'] = d

    # 9. The Gibbs emission-parameter draws: GaussianOutputModel.sample (gaussian.py:274-320) and
    #    DiscreteOutputModel.sample (discrete.py:217-251) consume numpy's global stream in a fixed order, so with
    #    np.random.seed they are reproducible: states with many, with one and with no observation.
    rng = np.random.default_rng(10)
    m0, s0 = np.array([-1.0, 0.5, 2.0, 7.0]), np.array([0.5, 1.0, 2.0, 0.3])
    obs_by_state = [m0[0] + s0[0] * rng.standard_normal(5000), m0[1] + s0[1] * rng.standard_normal(37),
                    np.array([2.25]), np.zeros(0)]
    gom = GaussianOutputModel(4, means=m0.copy(), sigmas=s0.copy())
    np.random.seed(123)
    gom.sample(obs_by_state)
    d = dict(means0=m0, sigmas0=s0, seed=np.int64(123), means=gom.means.copy(), sigmas=gom.sigmas.copy())
    for i, o in enumerate(obs_by_state):
        d['obs_in_state%d' % i] = o
    B0 = np.array([[0.5, 0.2, 0.1, 0.1, 0.05, 0.05], [0.05, 0.05, 0.1, 0.1, 0.2, 0.5], [0.2, 0.2, 0.2, 0.2, 0.1, 0.1]])
    sym_by_state = [rng.choice(6, size=800, p=B0[0]).astype(np.int32), np.array([5, 5, 4, 5], dtype=np.int32),
                    np.zeros(0, dtype=np.int32)]
    dom = DiscreteOutputModel(B0.copy())
    np.random.seed(77)
    dom.sample(sym_by_state)
    d.update(B0=B0, dseed=np.int64(77), B=dom.output_probabilities.copy())
    for i, o in enumerate(sym_by_state):
        d['sym_in_state%d' % i] = o
    out['gibbs_emission_draws'] = d

    # 10. The reference's own acceptance (bhmm/tests/test_hidden.py:240-256): impl 'c' np.allclose impl 'python'.  The numpy
    #     twin's outputs for the toy and the three-state example; the Viterbi path of the twin associates p (v A) instead of
    #     (p v) A (impl_python/hidden.py:207-247) and returns int64.
    d = {}
    hidden.set_implementation('python')
    for name in ('hidden_toy', 'hidden_gauss3'):
        g = out[name]
        lp, alpha = hidden.forward(g['A'], g['pobs'], g['pi'])
        beta = hidden.backward(g['A'], g['pobs'])
        gamma = hidden.state_probabilities(alpha, beta)
        d[name + '_logprob'] = np.float64(lp)
        d[name + '_alpha'], d[name + '_beta'], d[name + '_gamma'] = alpha, beta, gamma
        d[name + '_counts'] = hidden.state_counts(gamma, g['pobs'].shape[0])
        d[name + '_C'] = hidden.transition_counts(alpha, beta, g['A'], g['pobs'])
        d[name + '_viterbi'] = np.asarray(hidden.viterbi(g['A'], g['pobs'], g['pi']))
    hidden.set_implementation('c')
    out['hidden_python_twin'] = d

    only = [a.split('=', 1)[1] for a in sys.argv[2:] if a.startswith('--only=')]
    for name, arrays in out.items():
        if only and name not in only:
            continue
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **arrays)
        print('%-18s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    main()
