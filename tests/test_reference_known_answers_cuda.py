"""The reference's own known-answer tests for the estimator on pathological inputs (bhmm/tests/test_mlhmm_patho.py:27-72,
SURVEY 8c), run through bhmm_b200.estimate_hmm on the GPU engine with the reference's assertions and tolerances: tiny
discrete trajectories whose maximum-likelihood model is known in closed form (one state; a step; an alternation).  They
exercise what the large tests do not: a single trajectory of 3-9 frames, one hidden state, count matrices that are not
strongly connected under the default reversible=True (the estimator then takes the row-normalised counts between the
connected sets, _tmatrix_disconnected.py:68-190), and parameters that sit exactly on the boundary (0 and 1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _either(actual, ref, perm_ref, atol=1e-5):
    return np.allclose(actual, ref, atol=atol) or np.allclose(actual, perm_ref, atol=atol)


def test_1state():
    import bhmm_b200
    obs = np.array([0, 0, 0, 0, 0], dtype=int)
    hmm = bhmm_b200.estimate_hmm([obs], nstates=1, lag=1, accuracy=1e-6)
    assert np.allclose(hmm.initial_distribution, np.array([1.0]))
    assert np.allclose(hmm.transition_matrix, np.array([[1.0]]))
    assert np.allclose(hmm.output_model.output_probabilities, np.array([[1.0]]))


def test_1state_fail():
    """Two hidden states cannot be initialised from a trajectory that shows a single symbol (test_mlhmm_patho.py:39-42)."""
    import bhmm_b200
    obs = np.array([0, 0, 0, 0, 0], dtype=int)
    with pytest.raises(NotImplementedError):
        bhmm_b200.estimate_hmm([obs], nstates=2, lag=1, accuracy=1e-6)


def test_2state_step():
    import bhmm_b200
    obs = np.array([0, 0, 0, 0, 0, 1, 1, 1, 1], dtype=int)
    hmm = bhmm_b200.estimate_hmm([obs], nstates=2, lag=1, accuracy=1e-6)
    p0_ref = np.array([1.0, 0.0])
    A_ref = np.array([[0.8, 0.2], [0.0, 1.0]])
    B_ref = np.array([[1.0, 0.0], [0.0, 1.0]])
    perm = [1, 0]
    assert _either(hmm.initial_distribution, p0_ref, p0_ref[perm])
    assert _either(hmm.transition_matrix, A_ref, A_ref[np.ix_(perm, perm)])
    assert _either(hmm.output_model.output_probabilities, B_ref, B_ref[perm])


def test_2state_2step():
    import bhmm_b200
    obs = np.array([0, 1, 0], dtype=int)
    hmm = bhmm_b200.estimate_hmm([obs], nstates=2, lag=1, accuracy=1e-6)
    p0_ref = np.array([1.0, 0.0])
    A_ref = np.array([[0.0, 1.0], [1.0, 0.0]])
    B_ref = np.array([[1.0, 0.0], [0.0, 1.0]])
    perm = [1, 0]
    assert _either(hmm.initial_distribution, p0_ref, p0_ref[perm])
    assert _either(hmm.transition_matrix, A_ref, A_ref[np.ix_(perm, perm)])
    assert _either(hmm.output_model.output_probabilities, B_ref, B_ref[perm])


# ---- bhmm/tests/test_bhmm_patho.py:29-50: the Bayesian sampler on the same pathological inputs

def test_bayesian_2state_rev_step_refuses_disconnected_counts():
    import bhmm_b200
    obs = np.array([0, 0, 0, 0, 0, 1, 1, 1, 1], dtype=int)
    mle = bhmm_b200.estimate_hmm([obs], nstates=2, lag=1)
    with pytest.raises(NotImplementedError):      # disconnected count matrices with reversible sampling and no prior
        bhmm_b200.bayesian_hmm([obs], mle, reversible=True, p0_prior=None, transition_matrix_prior=None)


def test_bayesian_2state_nonrev_step():
    import bhmm_b200
    obs = np.array([0, 0, 0, 0, 0, 1, 1, 1, 1], dtype=int)
    mle = bhmm_b200.estimate_hmm([obs], nstates=2, lag=1)
    np.random.seed(20)            # the parameter draws use numpy's global stream like the reference: a fixed outcome
    sampled = bhmm_b200.bayesian_hmm([obs], mle, reversible=False, nsample=2000, p0_prior='mixed',
                                     transition_matrix_prior='mixed')
    # the absorbing state is whichever hidden state emits symbol 1 (the reference's test assumes it is state 1)
    absorbing = int(np.argmax(mle.output_model.output_probabilities[:, 1]))
    assert np.all(sampled.transition_matrix_std[1 - absorbing] > 0)
    assert np.max(np.abs(sampled.transition_matrix_std[absorbing])) < 1e-3


def test_bayesian_2state_rev_2step():
    import bhmm_b200
    obs = np.array([0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0], dtype=int)
    mle = bhmm_b200.estimate_hmm([obs], nstates=2, lag=1)
    np.random.seed(21)
    sampled = bhmm_b200.bayesian_hmm([obs], mle, reversible=False, nsample=100, p0_prior='mixed',
                                     transition_matrix_prior='mixed')
    assert np.all(sampled.transition_matrix_std > 0)
