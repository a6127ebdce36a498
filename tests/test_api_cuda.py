"""End to end through the reference's user entry points (bhmm/api.py:309-470) on the GPU engine: estimate_hmm without an
initial model (from-scratch heuristic + Baum-Welch), bayesian_hmm on top of it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sample(A, means, sigmas, lengths, seed):
    rng = np.random.default_rng(seed)
    N = len(means)
    cum = np.cumsum(A, axis=1)
    obs = []
    for T in lengths:
        s = np.empty(T, dtype=np.int64)
        s[0] = rng.integers(0, N)
        u = rng.random(T)
        for t in range(1, T):
            s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), N - 1)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
    return obs


def test_estimate_hmm_and_bayesian_hmm_gaussian():
    import bhmm_b200
    A = np.array([[0.97, 0.02, 0.01], [0.03, 0.94, 0.03], [0.01, 0.04, 0.95]])
    means, sigmas = np.array([-3.0, 0.0, 3.5]), np.array([0.8, 1.0, 0.6])
    obs = _sample(A, means, sigmas, [5000, 4000, 6000, 3000], 1)
    hmm = bhmm_b200.estimate_hmm(obs, 3, reversible=False, accuracy=1e-4, maxit=200)
    order = np.argsort(hmm.output_model.means)
    np.testing.assert_allclose(hmm.output_model.means[order], means, atol=0.1)
    np.testing.assert_allclose(hmm.output_model.sigmas[order], sigmas, rtol=0.1)
    np.testing.assert_allclose(hmm.transition_matrix[np.ix_(order, order)], A, atol=0.02)
    # lagged estimate: the transition matrix approximates A^2
    hmm2 = bhmm_b200.estimate_hmm(obs, 3, lag=2, reversible=False, accuracy=1e-4, maxit=200)
    o2 = np.argsort(hmm2.output_model.means)
    np.testing.assert_allclose(hmm2.transition_matrix[np.ix_(o2, o2)], A.dot(A), atol=0.03)
    assert hmm2._lag == 2
    post = bhmm_b200.bayesian_hmm(obs, hmm, nsample=6, reversible=False)
    assert post.nsamples == 6
    tm = post.transition_matrix
    assert tm['mean'].shape == (3, 3) and np.allclose(tm['samples'].sum(axis=2), 1.0)
    np.testing.assert_allclose(tm['mean'][np.ix_(order, order)], A, atol=0.03)
    np.testing.assert_allclose(post.means['mean'][order], means, atol=0.15)
    # reversible posterior (from-scratch sampler, parity unpinned): every sampled matrix is in detailed balance
    from bhmm_b200.util import tmatrix
    postr = bhmm_b200.bayesian_hmm(obs, hmm, nsample=3, reversible=True)
    for T in postr.transition_matrix['samples']:
        assert tmatrix.is_transition_matrix(T) and tmatrix.is_reversible(T)
    np.testing.assert_allclose(postr.transition_matrix['mean'][np.ix_(order, order)], A, atol=0.03)


def test_estimate_hmm_discrete():
    import bhmm_b200
    rng = np.random.default_rng(8)
    A = np.array([[0.96, 0.04], [0.05, 0.95]])
    B = np.array([[0.5, 0.3, 0.15, 0.05, 0.0, 0.0], [0.0, 0.02, 0.08, 0.2, 0.3, 0.4]])
    cum, cumB = np.cumsum(A, axis=1), np.cumsum(B, axis=1)
    obs = []
    for T in (8000, 6000):
        s = np.empty(T, dtype=np.int64)
        s[0] = 0
        u, v = rng.random(T), rng.random(T)
        for t in range(1, T):
            s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), 1)
        obs.append(np.array([min(int(np.searchsorted(cumB[k], x)), 5) for k, x in zip(s, v)], dtype=np.int64))
    hmm = bhmm_b200.estimate_hmm(obs, 2, reversible=False, accuracy=1e-4, maxit=300)
    Bh = hmm.output_model.output_probabilities
    order = np.argsort(Bh.dot(np.arange(6)))
    np.testing.assert_allclose(Bh[order], B, atol=0.03)
    np.testing.assert_allclose(hmm.transition_matrix[np.ix_(order, order)], A, atol=0.02)


def _golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name + '.npz'))


def test_estimate_hmm_matches_the_reference_entry_point_gaussian():
    """bhmm.estimate_hmm(observations, 3, lag=3, initial_model=..., reversible=False, accuracy=1e-3) of the REFERENCE
    (tests/golden/make_golden.py section 8, api.py:309-372): lagged observations, the estimator's own convergence test,
    fitted parameters to 1e-10, Viterbi paths of the nine lagged trajectories identical."""
    import bhmm_b200
    g = _golden('api_estimate_gauss3')
    obs = [g['obs%d' % k] for k in range(len(g['lengths']))]
    init = bhmm_b200.gaussian_hmm(g['pi0'], g['A0'], g['means0'], g['sigmas0'])
    hmm = bhmm_b200.estimate_hmm(obs, 3, lag=int(g['lag']), initial_model=init, reversible=False, stationary=False,
                                 accuracy=1e-3, maxit=200)
    assert hmm.lag == int(g['model_lag']) == 3
    assert abs(hmm.likelihood - float(g['likelihood'])) <= 1e-10 * abs(float(g['likelihood']))
    np.testing.assert_allclose(hmm.transition_matrix, g['A'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(hmm.initial_distribution, g['pi'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(hmm.output_model.means, g['means'], rtol=1e-10)
    np.testing.assert_allclose(hmm.output_model.sigmas, g['sigmas'], rtol=1e-10)
    paths = hmm.hidden_state_trajectories
    assert len(paths) == int(g['n_lagged']) == 9            # the 3-frame trajectory survives only at shift 0 (api.py:91)
    for k in range(len(paths)):
        assert np.array_equal(np.asarray(paths[k]), g['viterbi%d' % k])


def test_estimate_hmm_matches_the_reference_entry_point_discrete():
    import bhmm_b200
    g = _golden('api_estimate_discrete')
    obs = [g['obs%d' % k] for k in range(len(g['lengths']))]
    init = bhmm_b200.discrete_hmm(g['pi0'], g['A0'], g['B0'])
    hmm = bhmm_b200.estimate_hmm(obs, 3, lag=int(g['lag']), initial_model=init, reversible=False, stationary=False,
                                 accuracy=1e-3, maxit=200)
    assert hmm.lag == 2
    assert abs(hmm.likelihood - float(g['likelihood'])) <= 1e-10 * abs(float(g['likelihood']))
    np.testing.assert_allclose(hmm.transition_matrix, g['A'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(hmm.initial_distribution, g['pi'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(hmm.output_model.output_probabilities, g['B'], rtol=1e-10, atol=1e-14)
    for k in range(4):
        assert np.array_equal(np.asarray(hmm.hidden_state_trajectories[k]), g['viterbi%d' % k])
