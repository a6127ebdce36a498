"""Stand-in for the un-vendored, un-pinned `msmtools` dependency of bhmm (setup.py:118-122).

TEST INFRASTRUCTURE ONLY -- used by tests/golden/make_golden.py to import the reference package
in a container without msmtools and without network.  Only the calls that the NON-reversible
EM / Gibbs iteration reaches are provided; anything reversible raises.
"""
import sys
import types

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components


def is_transition_matrix(T, tol=1e-12):
    T = np.asarray(T)
    return T.ndim == 2 and T.shape[0] == T.shape[1] and np.all(T >= -tol) and np.allclose(T.sum(axis=1), 1.0, atol=1e-8)


def stationary_distribution(T, **kw):
    T = np.asarray(T, dtype=float)
    w, v = np.linalg.eig(T.T)
    k = np.argmax(w.real)
    p = np.abs(v[:, k].real)
    return p / p.sum()


def connected_sets(C, directed=True):
    n, labels = connected_components(sp.csr_matrix(np.asarray(C) > 0), directed=True,
                                     connection='strong' if directed else 'weak')
    sets = [np.where(labels == k)[0] for k in range(n)]
    sets.sort(key=lambda s: -len(s))
    return sets


def largest_connected_set(C, directed=True):
    return connected_sets(C, directed)[0]


def transition_matrix(C, reversible=False, **kw):
    if reversible:
        raise NotImplementedError("msmtools stub: reversible estimator is not pinned (SURVEY.md 8c)")
    C = np.asarray(C, dtype=float)
    return C / C.sum(axis=1)[:, None]


def count_matrix(dtrajs, lag, nstates=None, **kw):
    if isinstance(dtrajs, np.ndarray) and dtrajs.ndim == 1:
        dtrajs = [dtrajs]
    if nstates is None:
        nstates = max(int(np.max(d)) for d in dtrajs) + 1
    Cm = np.zeros((nstates, nstates))
    for d in dtrajs:
        d = np.asarray(d)
        np.add.at(Cm, (d[:-lag], d[lag:]), 1.0)
    return sp.csr_matrix(Cm)


def sample_tmatrix(C, nsample=1, nsteps=None, reversible=False, **kw):
    if reversible:
        raise NotImplementedError("msmtools stub: reversible sampler is not pinned (SURVEY.md 8c)")
    C = np.asarray(C, dtype=float)
    P = np.zeros_like(C)
    for i in range(C.shape[0]):
        pos = C[i] > 0
        P[i, pos] = np.random.dirichlet(C[i, pos])
    return P


def generate_traj(P, N, start=None, stop=None, dt=1):
    P = np.asarray(P)
    n = P.shape[0]
    s = np.zeros(N, dtype=int)
    s[0] = np.random.choice(n, p=stationary_distribution(P)) if start is None else start
    cum = np.cumsum(P, axis=1)
    u = np.random.random(N)
    for t in range(1, N):
        s[t] = min(np.searchsorted(cum[s[t - 1]], u[t]), n - 1)
    return s


def install():
    """Register the stand-in as `msmtools` (and the sub-modules bhmm imports) in sys.modules."""
    root = types.ModuleType('msmtools')
    ana = types.ModuleType('msmtools.analysis')
    est = types.ModuleType('msmtools.estimation')
    gen = types.ModuleType('msmtools.generation')
    dense = types.ModuleType('msmtools.analysis.dense')
    sv = types.ModuleType('msmtools.analysis.dense.stationary_vector')
    ana.is_transition_matrix = is_transition_matrix
    ana.stationary_distribution = stationary_distribution
    sv.stationary_distribution = stationary_distribution
    est.connected_sets = connected_sets
    est.largest_connected_set = largest_connected_set
    est.transition_matrix = transition_matrix
    est.count_matrix = count_matrix
    est.sample_tmatrix = sample_tmatrix
    gen.generate_traj = generate_traj
    root.analysis, root.estimation, root.generation = ana, est, gen
    ana.dense, dense.stationary_vector = dense, sv
    for m in (root, ana, est, gen, dense, sv):
        sys.modules[m.__name__] = m
