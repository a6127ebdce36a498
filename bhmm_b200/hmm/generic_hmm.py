"""Parameter container with the interface of bhmm/hmm/generic_hmm.py:25-431 (the parts the estimators use)."""
import numpy as np

from ..util import tmatrix as _tmatrix


class HMM(object):
    r""" Hidden Markov model (HMM): initial distribution, hidden transition matrix and an output model
    (bhmm/hmm/generic_hmm.py:25-78)."""

    def __init__(self, Pi, Tij, output_model, lag=1):
        self._nstates = np.shape(Tij)[0]
        self._lag = lag
        self.output_model = output_model
        self.hidden_state_trajectories = None
        self.likelihood = None
        self.update(Pi, Tij)

    def update(self, Pi, Tij):
        r""" Updates the transition matrix and recomputes all derived quantities (generic_hmm.py:79-93) """
        self._Tij = np.array(Tij)
        assert _tmatrix.is_transition_matrix(self._Tij), 'Given transition matrix is not a stochastic matrix'
        assert self._Tij.shape[0] == self._nstates, 'Given transition matrix has unexpected number of states '
        Pi = np.asarray(Pi, dtype=float)
        assert np.all(Pi >= 0), 'Given initial distribution contains negative elements.'
        assert np.any(Pi > 0), 'Given initial distribution is zero'
        self._Pi = np.array(Pi) / np.sum(Pi)

    def __repr__(self):
        return "HMM(%d, %s, %s, Pi=%s, stationary=%s, reversible=%s)" % (
            self._nstates, repr(self._Tij), repr(self.output_model), repr(self._Pi), repr(self.is_stationary),
            repr(self.is_reversible))

    @property
    def lag(self):
        return self._lag

    @property
    def nstates(self):
        return self._nstates

    @property
    def initial_distribution(self):
        return self._Pi

    @property
    def transition_matrix(self):
        return self._Tij

    @property
    def is_reversible(self):
        return _tmatrix.is_reversible(self._Tij)

    @property
    def is_stationary(self):
        return np.allclose(np.dot(self._Pi, self._Tij), self._Pi)

    @property
    def stationary_distribution(self):
        assert _tmatrix.is_connected(self._Tij, strong=False), \
            'No unique stationary distribution because transition matrix is not connected'
        return _tmatrix.stationary_vector(self._Tij)

    @property
    def lifetimes(self):
        return -self._lag / np.log(np.diag(self.transition_matrix))

    # ---- Gibbs path statistics on the host (the engine computes the same numbers on the GPU)
    def count_matrix(self):
        """Lag-1 transition counts of the hidden state trajectories (generic_hmm.py:297-319)."""
        if self.hidden_state_trajectories is None:
            raise RuntimeError('HMM model does not have a hidden state trajectory.')
        Cm = np.zeros((self._nstates, self._nstates))
        for s in self.hidden_state_trajectories:
            s = np.asarray(s)
            np.add.at(Cm, (s[:-1], s[1:]), 1.0)
        return Cm

    def count_init(self):
        """Counts at the first time step (generic_hmm.py:321-334)."""
        if self.hidden_state_trajectories is None:
            raise RuntimeError('HMM model does not have a hidden state trajectory.')
        n = [traj[0] for traj in self.hidden_state_trajectories]
        return np.bincount(n, minlength=self.nstates)

    def collect_observations_in_state(self, observations, state_index):
        """All observations assigned to one hidden state (generic_hmm.py:398-431)."""
        if not self.hidden_state_trajectories:
            raise RuntimeError('HMM model does not have a hidden state trajectory.')
        parts = [np.asarray(o)[np.where(np.asarray(s) == state_index)[0]]
                 for s, o in zip(self.hidden_state_trajectories, observations)]
        return np.concatenate(parts) if parts else np.array([])
