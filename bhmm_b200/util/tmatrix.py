"""Transition-matrix M-step helpers (host, N x N): the part of bhmm/estimators/_tmatrix_disconnected.py the EM and
Gibbs iterations reach, written without msmtools (which the reference imports for every one of these).

Parity status: the NON-reversible estimator (C / rowsum on weakly connected sets, _tmatrix_disconnected.py:106-115)
is pinned by tests/golden/em_*.npz.  The reversible estimator is a from-scratch fixed-point iteration of the
detailed-balance MLE; the reference delegates it to an absent, un-pinned msmtools, so it is PARITY-UNPINNED
 and tested only through its defining properties.
"""
import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components


def connected_sets(C, mincount_connectivity=0, strong=True):
    """Connected sets of the count graph, largest first (_tmatrix_disconnected.py:27-43)."""
    Cconn = np.array(C, dtype=float)
    Cconn[np.where(Cconn <= mincount_connectivity)] = 0
    n, labels = connected_components(csr_matrix(Cconn > 0), directed=True, connection='strong' if strong else 'weak')
    sets = [np.where(labels == k)[0] for k in range(n)]
    sets.sort(key=lambda s: -len(s))
    return sets


def is_connected(C, mincount_connectivity=0, strong=True):
    return len(connected_sets(C, mincount_connectivity=mincount_connectivity, strong=strong)) == 1


def is_transition_matrix(T, tol=1e-10):
    T = np.asarray(T)
    return bool(T.ndim == 2 and T.shape[0] == T.shape[1] and np.all(T >= -tol)
                and np.allclose(T.sum(axis=1), 1.0, atol=1e-8))


def stationary_vector(P):
    """Stationary distribution of a connected stochastic matrix (dense eigen-decomposition)."""
    P = np.asarray(P, dtype=float)
    if P.shape[0] == 1:
        return np.ones(1)
    w, v = np.linalg.eig(P.T)
    k = np.argmax(w.real)
    mu = np.abs(v[:, k].real)
    return mu / mu.sum()


def stationary_distribution(P, C=None, mincount_connectivity=0):
    """Stationary distribution, weighting disconnected sets by their counts (_tmatrix_disconnected.py:229-251)."""
    if C is None:
        if is_connected(P, strong=True):
            return stationary_vector(P)
        raise ValueError('Computing stationary distribution for disconnected matrix. Need count matrix.')
    n = np.shape(C)[0]
    ctot = np.sum(C)
    pi = np.zeros(n)
    for s in connected_sets(C, mincount_connectivity=mincount_connectivity, strong=False):
        w = np.sum(C[s, :]) / ctot
        pi[s] = w * stationary_vector(P[s, :][:, s])
    return pi / np.sum(pi)


def is_reversible(P):
    """Detailed balance on every weakly connected set (_tmatrix_disconnected.py:211-226)."""
    P = np.asarray(P, dtype=float)
    for s in connected_sets(P, strong=False):
        Ps = P[s, :][:, s]
        if not is_transition_matrix(Ps):
            return False
        pi = stationary_vector(Ps)
        X = pi[:, None] * Ps
        if not np.allclose(X, X.T):
            return False
    return True


def transition_matrix_reversible(C, maxiter=1000000, maxerr=1e-12):
    """Reversible maximum-likelihood transition matrix of a strongly connected count matrix by the classical
    fixed-point iteration x_ij <- (c_ij + c_ji) / (c_i / x_i + c_j / x_j).  PARITY-UNPINNED (see module doc)."""
    C = np.asarray(C, dtype=float)
    Csym = C + C.T
    ci = C.sum(axis=1)
    X = Csym.copy()
    X /= X.sum()
    pi_old = X.sum(axis=1)
    for _ in range(int(maxiter)):
        xi = X.sum(axis=1)
        q = ci / xi
        denom = q[:, None] + q[None, :]
        X = np.where(denom > 0, Csym / np.where(denom > 0, denom, 1.0), 0.0)
        X /= X.sum()
        pi_new = X.sum(axis=1)
        if np.max(np.abs(pi_new - pi_old)) < maxerr:
            break
        pi_old = pi_new
    return X / X.sum(axis=1)[:, None]


def estimate_P(C, reversible=True, fixed_statdist=None, maxiter=1000000, maxerr=1e-8, mincount_connectivity=0):
    """Full transition matrix for general connectivity (_tmatrix_disconnected.py:68-123)."""
    C = np.array(C, dtype=float)
    n = C.shape[0]
    if not reversible and fixed_statdist is None and np.all(C > mincount_connectivity):
        return C / C.sum(axis=1)[:, None]      # one connected set, no empty row: the general code below reduces to this
    P = np.eye(n, dtype=np.float64)
    if fixed_statdist is not None:
        raise NotImplementedError('estimation with a fixed stationary distribution is not part of the hot path')
    if reversible:
        for s in connected_sets(C, mincount_connectivity=mincount_connectivity, strong=True):
            mask = np.zeros(n, dtype=bool)
            mask[s] = True
            if C[np.ix_(mask, ~mask)].sum() > np.finfo(C.dtype).eps:
                raise NotImplementedError('partially reversible estimation (transient sets) is not implemented; '
                                          'use reversible=False')
            if s.size > 1:
                I = np.ix_(mask, mask)
                P[I] = transition_matrix_reversible(C[I], maxiter=maxiter, maxerr=maxerr)
    else:
        for s in connected_sets(C, mincount_connectivity=mincount_connectivity, strong=False):
            I = np.ix_(s, s)
            Csub = C[I]
            zero_rows = np.where(Csub.sum(axis=1) == 0)[0]
            Csub[zero_rows, zero_rows] = 1.0
            P[I] = Csub / Csub.sum(axis=1)[:, None]
    return P


def sample_P_reversible(C, nsteps=1000, P0=None, rng=None):
    """One draw from the posterior of REVERSIBLE transition matrices given the (posterior) count matrix C.

    From scratch -- the reference delegates this to msmtools.estimation.sample_tmatrix (bayesian_sampling.py:355-356),
    which is absent and un-pinned, so this sampler is PARITY-UNPINNED: it is tested by its defining properties only
    (samples are stochastic and obey detailed balance, their mean approaches the reversible maximum-likelihood
    estimate as counts grow, their spread shrinks like 1/sqrt(counts)).

    Parametrisation: a symmetric non-negative weight matrix X with T_ij = x_ij / sum_k x_ik, which is reversible with
    stationary vector proportional to the row sums of X.  Target: prod_ij T_ij^(c_ij) with respect to the flat measure on
    the NORMALISED weights (the simplex sum_{i<=j} x_ij = 1) -- c_ij are the caller's posterior counts, i.e. observed
    counts plus prior counts, exactly what the reference hands to its sampler.  T does not depend on the scale of X, so
    the flat measure on the unnormalised weights would be improper (round-1 advisor finding: a chain that is merely
    renormalised after every sweep samples a 1/x-type prior instead); the scale is therefore given an independent
    Exp(1) weight per x_ij -- normalised independent Gamma(1, 1) variables are uniform on the simplex and independent of
    their sum -- which makes the target proper without changing the law of T.  `nsteps` Metropolis sweeps over the
    weights on the sparsity pattern of C + C^T, log-normal random-walk proposals (the Jacobian of the log
    parametrisation is part of the acceptance ratio), started at the reversible maximum-likelihood estimate (or P0)
    scaled to the prior mean of the total weight.
    """
    rng = np.random.default_rng() if rng is None else rng
    C = np.asarray(C, dtype=np.float64)
    n = C.shape[0]
    if not is_connected(C, strong=True):
        raise NotImplementedError('Encountered disconnected count matrix with sampling option reversible:\n ' + str(C)
                                  + '\nUse prior to ensure connectivity or use reversible=False.')
    if P0 is None:
        P0 = transition_matrix_reversible(C)
    pi = stationary_vector(P0)
    X = pi[:, None] * P0
    X = 0.5 * (X + X.T)
    Csym = C + C.T
    pattern = [(i, j) for i in range(n) for j in range(i, n) if Csym[i, j] > 0 and X[i, j] > 0]
    X *= len(pattern) / sum(X[i, j] for (i, j) in pattern)     # total weight at its prior mean (one Exp(1) per weight)
    rows = X.sum(axis=1)
    crow = C.sum(axis=1)
    sigma = 0.5
    for _ in range(int(nsteps)):
        for (i, j) in pattern:
            old = X[i, j]
            new = old * np.exp(sigma * rng.standard_normal())
            d = new - old
            if i == j:
                ri = rows[i] + d
                # log f = sum c_ab log x_ab - sum_a crow_a log rows_a ; only x_ii and rows_i change
                dlog = C[i, i] * np.log(new / old) - crow[i] * np.log(ri / rows[i])
            else:
                ri, rj = rows[i] + d, rows[j] + d
                dlog = (C[i, j] + C[j, i]) * np.log(new / old) - crow[i] * np.log(ri / rows[i]) \
                    - crow[j] * np.log(rj / rows[j])
            dlog += np.log(new / old) - d                # Jacobian of the log-normal proposal; Exp(1) weight of x_ij
            if np.log(rng.random()) < dlog:
                X[i, j] = new
                if i == j:
                    rows[i] = ri
                else:
                    X[j, i] = new
                    rows[i], rows[j] = ri, rj
        rows = X.sum(axis=1)                             # (recomputed once per sweep: no drift of the running sums)
    return X / rows[:, None]
