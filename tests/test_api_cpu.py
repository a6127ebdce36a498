"""Host-side callers of the hot path (bhmm_b200.api; reference bhmm/api.py:23-306): lagging, output-type guess and the
from-scratch initial-model heuristics.  No GPU needed."""
import numpy as np
import pytest

from bhmm_b200 import api
from bhmm_b200.util import testsystems as ts


def test_lag_observations_matches_reference_rule():
    obs = [np.arange(10), np.arange(100, 104), np.arange(3)]
    out = api.lag_observations(obs, 3)
    # (s0,s3,s6,s9), (s1,s4,s7), (s2,s5,s8) / (100,103); (101,) and (102,) dropped / third trajectory: one frame each
    assert [list(o) for o in out] == [[0, 3, 6, 9], [1, 4, 7], [2, 5, 8], [100, 103]]
    assert [list(o) for o in api.lag_observations(obs, 3, stride=3)] == [[0, 3, 6, 9], [100, 103]]
    assert [list(o) for o in api.lag_observations([np.arange(4)], 1)] == [[0, 1, 2, 3]]


def test_lag_observations_matches_the_reference_fixture():
    """bhmm.lag_observations of the reference on four ragged trajectories (tests/golden/make_golden.py section 8): number and
    lengths of the lagged trajectories at lag 3, and lengths / first / last frames at lag 4 with stride 2."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'api_estimate_gauss3.npz'))
    obs = [g['obs%d' % k] for k in range(len(g['lengths']))]
    out = api.lag_observations(obs, 3)
    assert len(out) == int(g['n_lagged']) and [len(o) for o in out] == list(g['lagged_lengths'])
    st = api.lag_observations(obs, 4, stride=2)
    assert [len(o) for o in st] == list(g['strided_lengths'])
    assert np.array_equal([o[0] for o in st], g['strided_heads']) and np.array_equal([o[-1] for o in st], g['strided_tails'])


def test_guess_output_type():
    assert api._guess_output_type([np.array([0, 1, 2])]) == 'discrete'
    assert api._guess_output_type([np.array([0.0, 1.0, 2.0])]) == 'discrete'
    assert api._guess_output_type([np.array([0.1, 1.0, 2.0])]) == 'gaussian'
    with pytest.raises(TypeError):
        api._guess_output_type([np.zeros((3, 2))])


def _sample_gaussian(A, means, sigmas, lengths, seed):
    rng = np.random.default_rng(seed)
    N = len(means)
    cum = np.cumsum(A, axis=1)
    obs, paths = [], []
    for T in lengths:
        s = np.empty(T, dtype=np.int64)
        s[0] = rng.integers(0, N)
        u = rng.random(T)
        for t in range(1, T):
            s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), N - 1)
        obs.append(means[s] + sigmas[s] * rng.standard_normal(T))
        paths.append(s)
    return obs, paths


def test_init_gaussian_hmm_recovers_well_separated_states():
    A = np.array([[0.95, 0.04, 0.01], [0.05, 0.9, 0.05], [0.02, 0.08, 0.9]])
    means, sigmas = np.array([-4.0, 0.0, 5.0]), np.array([0.7, 1.0, 0.5])
    obs, _ = _sample_gaussian(A, means, sigmas, [4000, 2500, 3000], 5)
    m = api.init_gaussian_hmm(obs, 3, reversible=False)
    np.testing.assert_allclose(m.output_model.means, means, atol=0.15)
    np.testing.assert_allclose(m.output_model.sigmas, sigmas, rtol=0.15)
    P = m.transition_matrix
    assert np.all(P >= 0) and np.allclose(P.sum(axis=1), 1.0)
    np.testing.assert_allclose(P, A, atol=0.03)
    assert abs(m.initial_distribution.sum() - 1.0) < 1e-12 and np.all(m.initial_distribution > 0)
    # the reversible variant satisfies detailed balance
    mr = api.init_gaussian_hmm(obs, 3, reversible=True)
    Pr = mr.transition_matrix
    from bhmm_b200.util import tmatrix
    pi = tmatrix.stationary_distribution(Pr)
    np.testing.assert_allclose(pi[:, None] * Pr, (pi[:, None] * Pr).T, atol=1e-8)


def test_init_discrete_hmm_finds_the_metastable_lumping():
    rng = np.random.default_rng(3)
    # 3 hidden states, each emitting its own block of 4 symbols
    A = np.array([[0.97, 0.02, 0.01], [0.02, 0.96, 0.02], [0.01, 0.03, 0.96]])
    cum = np.cumsum(A, axis=1)
    obs = []
    for T in (6000, 5000):
        s = np.empty(T, dtype=np.int64)
        s[0] = rng.integers(0, 3)
        u = rng.random(T)
        for t in range(1, T):
            s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), 2)
        obs.append((4 * s + rng.integers(0, 4, size=T)).astype(np.int32))
    m = api.init_discrete_hmm(obs, 3, reversible=True)
    B = m.output_model.output_probabilities
    assert B.shape == (3, 12) and np.allclose(B.sum(axis=1), 1.0)
    # every hidden state owns one block of symbols (up to a permutation of the states)
    owner = np.argmax(B, axis=0)
    blocks = [set(owner[4 * k:4 * k + 4]) for k in range(3)]
    assert all(len(b) == 1 for b in blocks) and len(set.union(*blocks)) == 3
    P = m.transition_matrix
    assert np.allclose(P.sum(axis=1), 1.0) and np.all(np.diag(P) > 0.9)
    with pytest.raises(NotImplementedError):      # like the reference: init/discrete.py:270-272, test_mlhmm_patho.py:39-42
        api.init_discrete_hmm([np.array([0, 1, 0, 1])], 3)


def test_model_constructors():
    g = api.gaussian_hmm([0.5, 0.5], [[0.9, 0.1], [0.2, 0.8]], [-1.0, 1.0], [0.5, 0.6])
    assert g.nstates == 2 and g.output_model.model_type == 'gaussian'
    d = api.discrete_hmm([0.5, 0.5], [[0.9, 0.1], [0.2, 0.8]], [[0.7, 0.3, 0.0], [0.1, 0.1, 0.8]])
    assert d.nstates == 2 and d.output_model.model_type == 'discrete'
