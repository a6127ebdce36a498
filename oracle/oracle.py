"""oracle/oracle.py -- ctypes front-end of the CPU checker.  TEST INFRASTRUCTURE ONLY.

Two libraries can sit behind the same Python surface:

* ``liboracle.so``            our plain-C restatement (oracle/hmm_oracle.c), always available;
* ``_ref/libbhmm_ref.so``     the reference's own C sources (bhmm/hidden/impl_c/_hidden.c,
                              bhmm/output_models/impl_c/_gaussian.c, _discrete.c) compiled in
                              place by ``make -C oracle ref``; present wherever /root/reference
                              was available at build time (it travels to the GPU box prebuilt).

``Oracle(kind='port')`` and ``Oracle(kind='reference')`` expose the same methods with the
reference's array conventions (C-contiguous float64, (T,N) time-major, int32 paths), so a test
can run either against the CUDA path, and tests/test_oracle.py can pin one against the other.

The numpy-level steps that the reference performs in Python are restated here as well:
``state_probabilities`` / ``state_counts`` (bhmm/hidden/api.py:133-211), the outlier rule
(bhmm/output_models/outputmodel.py:119-131), the Gaussian / discrete M-steps
(bhmm/output_models/gaussian.py:214-272, discrete.py:159-215) and the non-reversible
transition-matrix update (bhmm/estimators/_tmatrix_disconnected.py:110-115 -> C / rowsum).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _l(a):
    return a.ctypes.data_as(_lp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(ref=True):
    """Compile liboracle.so (and _ref/libbhmm_ref.so when the reference checkout exists)."""
    targets = ["all"] + (["ref"] if ref else [])
    # make's chatter must not reach stdout (bench.py prints exactly one JSON line there)
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True, stdout=subprocess.DEVNULL)


def have_reference_lib():
    return os.path.exists(os.path.join(_HERE, "_ref", "libbhmm_ref.so"))


class Oracle(object):
    """CPU implementation of the hot path; ``kind`` is 'port' (restatement) or 'reference'."""

    def __init__(self, kind="port"):
        self.kind = kind
        port_path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(port_path):
            build(ref=False)
        self._port = C.CDLL(port_path)
        self._ref = None
        if kind == "reference":
            ref_path = os.path.join(_HERE, "_ref", "libbhmm_ref.so")
            if not os.path.exists(ref_path):
                raise RuntimeError("oracle/_ref/libbhmm_ref.so missing: run `make -C oracle ref` "
                                   "where /root/reference is available")
            self._ref = C.CDLL(ref_path)
            self._ref._forward.restype = C.c_double
            self._ref._backward.restype = None
            self._ref._p_obs.restype = None
            self._ref._update_pout.restype = None
            self._ref.set_seed.restype = None
        elif kind != "port":
            raise ValueError(kind)
        p = self._port
        p.orc_forward.restype = C.c_double
        p.orc_estep_gaussian_traj.restype = C.c_double
        p.orc_estep_discrete_traj.restype = C.c_double
        for name in ("orc_backward", "orc_gaussian_pobs", "orc_discrete_pobs", "orc_state_probabilities",
                     "orc_state_counts", "orc_discrete_update_pout", "orc_glibc_uniforms",
                     "orc_gaussian_var_pass", "orc_path_stats"):
            getattr(p, name).restype = None

    # ---------------------------------------------------------------- emission
    def gaussian_p_obs(self, obs, means, sigmas, ignore_outliers=True):
        """GaussianOutputModel.p_obs (gaussian.py:170-212) incl. the outlier rule."""
        obs, means, sigmas = _f64(obs), _f64(means), _f64(sigmas)
        T, N = obs.shape[0], means.shape[0]
        out = np.zeros((T, N))
        if self._ref is not None:
            self._ref._p_obs(_d(obs), _d(means), _d(sigmas), C.c_int(N), C.c_int(T), _d(out))
        else:
            self._port.orc_gaussian_pobs(_d(obs), _d(means), _d(sigmas), N, T, _d(out))
        if ignore_outliers:
            self._port.orc_handle_outliers(_d(out), N, T)
        return out

    def discrete_p_obs(self, obs, B, ignore_outliers=False):
        """DiscreteOutputModel.p_obs (discrete.py:130-157)."""
        obs = np.ascontiguousarray(obs, dtype=np.int32)
        B = _f64(B)
        N, M = B.shape
        out = np.zeros((obs.shape[0], N))
        self._port.orc_discrete_pobs(_i(obs), _d(B), N, M, obs.shape[0], _d(out))
        if ignore_outliers:
            self._port.orc_handle_outliers(_d(out), N, obs.shape[0])
        return out

    def update_pout(self, obs, weights, pout):
        """_update_pout scatter-add (impl_c/_discrete.c:1-32); pout (N,M) is updated in place."""
        obs = np.ascontiguousarray(obs, dtype=np.int32)
        weights = _f64(weights)
        N, M = pout.shape
        fn = self._ref._update_pout if self._ref is not None else self._port.orc_discrete_update_pout
        fn(_i(obs), _d(weights), C.c_int(obs.shape[0]), C.c_int(N), C.c_int(M), _d(pout))
        return pout

    # ---------------------------------------------------------------- hidden
    def forward(self, A, pobs, pi, T=None):
        A, pobs, pi = _f64(A), _f64(pobs), _f64(pi)
        T = pobs.shape[0] if T is None else T
        N = A.shape[0]
        alpha = np.zeros((T, N))
        if self._ref is not None:
            lp = self._ref._forward(_d(alpha), _d(A), _d(pobs), _d(pi), C.c_int(N), C.c_int(T))
        else:
            lp = self._port.orc_forward(_d(alpha), _d(A), _d(pobs), _d(pi), N, T)
        return lp, alpha

    def backward(self, A, pobs, T=None):
        A, pobs = _f64(A), _f64(pobs)
        T = pobs.shape[0] if T is None else T
        N = A.shape[0]
        beta = np.zeros((T, N))
        if self._ref is not None:
            self._ref._backward(_d(beta), _d(A), _d(pobs), C.c_int(N), C.c_int(T))
        else:
            self._port.orc_backward(_d(beta), _d(A), _d(pobs), N, T)
        return beta

    def state_probabilities(self, alpha, beta):
        alpha, beta = _f64(alpha), _f64(beta)
        T, N = alpha.shape
        gamma = np.zeros((T, N))
        self._port.orc_state_probabilities(_d(gamma), _d(alpha), _d(beta), N, T)
        return gamma

    def state_counts(self, gamma, T=None):
        gamma = _f64(gamma)
        T = gamma.shape[0] if T is None else T
        out = np.zeros(gamma.shape[1])
        self._port.orc_state_counts(_d(out), _d(gamma), gamma.shape[1], T)
        return out

    def transition_counts(self, alpha, beta, A, pobs, T=None):
        alpha, beta, A, pobs = _f64(alpha), _f64(beta), _f64(A), _f64(pobs)
        T = pobs.shape[0] if T is None else T
        N = A.shape[0]
        Cm = np.zeros((N, N))
        if self._ref is not None:
            rc = self._ref._compute_transition_counts(_d(Cm), _d(A), _d(pobs), _d(alpha), _d(beta),
                                                      C.c_int(N), C.c_int(T))
        else:
            rc = self._port.orc_transition_counts(_d(Cm), _d(A), _d(pobs), _d(alpha), _d(beta), N, T)
        if rc:
            raise MemoryError()
        return Cm

    def viterbi(self, A, pobs, pi):
        A, pobs, pi = _f64(A), _f64(pobs), _f64(pi)
        T, N = pobs.shape
        path = np.zeros(T, dtype=np.int32)
        if self._ref is not None:
            rc = self._ref._compute_viterbi(_i(path), _d(A), _d(pobs), _d(pi), C.c_int(N), C.c_int(T))
        else:
            rc = self._port.orc_viterbi(_i(path), _d(A), _d(pobs), _d(pi), N, T)
        if rc:
            raise MemoryError()
        return path

    def glibc_uniforms(self, seed, n):
        """The uniforms r = rand()/(RAND_MAX+1.0) that follow srand(seed) in glibc, in draw order."""
        u = np.zeros(n)
        self._port.orc_glibc_uniforms(C.c_int(int(seed)), C.c_long(n), _d(u))
        return u

    def sample_path(self, alpha, A, T=None, seed=None, u=None):
        """_sample_path (_hidden.c:330-378).  Either explicit uniforms ``u`` (draw order: first for
        t=T-1) or a glibc ``seed``.  With kind='reference' and a seed, the reference's own
        set_seed()/rand() path runs (libc state)."""
        alpha, A = _f64(alpha), _f64(A)
        T = alpha.shape[0] if T is None else T
        N = A.shape[0]
        path = np.zeros(T, dtype=np.int32)
        if self._ref is not None and u is None:
            self._ref.set_seed(C.c_int(int(seed)))
            dummy = np.zeros((1, N))
            rc = self._ref._sample_path(_i(path), _d(alpha), _d(A), _d(dummy), C.c_int(N), C.c_int(T))
        else:
            if u is None:
                u = self.glibc_uniforms(seed, T)
            u = _f64(u)
            rc = self._port.orc_sample_path(_i(path), _d(alpha), _d(A), _d(u), N, T)
        if rc:
            raise RuntimeError("sample_path failed with code %d" % rc)
        return path

    # ---------------------------------------------------------------- composites
    def estep_gaussian(self, observations, A, pi, means, sigmas, ignore_outliers=True):
        """E-step over a list of trajectories in the reference's call sequence
        (maximum_likelihood.py:383-385 + :271-282).  Uses the reference library for the five C
        calls when kind='reference'.  Returns dict(loglik, gamma0, C, wsum, wo, gammas)."""
        A, pi, means, sigmas = _f64(A), _f64(pi), _f64(means), _f64(sigmas)
        N = A.shape[0]
        g0, Cm, ws, wo = np.zeros(N), np.zeros((N, N)), np.zeros(N), np.zeros(N)
        ll, gammas = 0.0, []
        for obs in observations:
            obs = _f64(obs)
            T = obs.shape[0]
            if self._ref is None:
                work = np.zeros(3 * T * N + N * N)
                gamma = np.zeros((T, N))
                ll += self._port.orc_estep_gaussian_traj(_d(obs), T, N, _d(A), _d(pi), _d(means), _d(sigmas),
                                                         int(bool(ignore_outliers)), _d(work), _d(gamma),
                                                         _d(g0), _d(Cm), _d(ws), _d(wo))
            else:
                pobs = self.gaussian_p_obs(obs, means, sigmas, ignore_outliers)
                lp, alpha = self.forward(A, pobs, pi)
                beta = self.backward(A, pobs)
                gamma = self.state_probabilities(alpha, beta)
                Cm += self.transition_counts(alpha, beta, A, pobs)
                g0 += gamma[0]
                ws += gamma.sum(axis=0)
                wo += gamma.T.dot(obs)
                ll += lp
            gammas.append(gamma)
        return dict(loglik=ll, gamma0=g0, C=Cm, wsum=ws, wo=wo, gammas=gammas)

    def estep_discrete(self, observations, A, pi, B, ignore_outliers=False):
        A, pi, B = _f64(A), _f64(pi), _f64(B)
        N, M = B.shape
        g0, Cm, Bnum = np.zeros(N), np.zeros((N, N)), np.zeros((N, M))
        ll, gammas = 0.0, []
        for obs in observations:
            obs = np.ascontiguousarray(obs, dtype=np.int32)
            T = obs.shape[0]
            if self._ref is None:
                work = np.zeros(3 * T * N + N * N)
                gamma = np.zeros((T, N))
                ll += self._port.orc_estep_discrete_traj(_i(obs), T, N, M, _d(A), _d(pi), _d(B),
                                                         int(bool(ignore_outliers)), _d(work), _d(gamma),
                                                         _d(g0), _d(Cm), _d(Bnum))
            else:
                pobs = self.discrete_p_obs(obs, B, ignore_outliers)
                lp, alpha = self.forward(A, pobs, pi)
                beta = self.backward(A, pobs)
                gamma = self.state_probabilities(alpha, beta)
                Cm += self.transition_counts(alpha, beta, A, pobs)
                g0 += gamma[0]
                self.update_pout(obs, gamma, Bnum)
                ll += lp
            gammas.append(gamma)
        return dict(loglik=ll, gamma0=g0, C=Cm, Bnum=Bnum, gammas=gammas)

    def path_stats(self, paths, observations, N):
        """Gibbs path statistics (generic_hmm.py:297-334,398-431): integer lag-1 count matrix,
        first-state histogram, per-state frame counts, sum(o), sum(o^2)."""
        Cint = np.zeros((N, N), dtype=np.int64)
        n0 = np.zeros(N, dtype=np.int64)
        cnt = np.zeros(N, dtype=np.int64)
        so, soo = np.zeros(N), np.zeros(N)
        for path, obs in zip(paths, observations):
            path = np.ascontiguousarray(path, dtype=np.int32)
            obs = _f64(obs)
            self._port.orc_path_stats(_i(path), _d(obs), path.shape[0], N, _l(Cint), _l(n0), _l(cnt),
                                      _d(so), _d(soo))
        return dict(C=Cint, n0=n0, count=cnt, so=so, soo=soo)


# ------------------------------------------------------------------------- M-step (numpy, host)
def mstep_gaussian(observations, gammas):
    """GaussianOutputModel.estimate (gaussian.py:214-272): two passes, variance about the NEW mean."""
    N = gammas[0].shape[1]
    means, wsum = np.zeros(N), np.zeros(N)
    for o, g in zip(observations, gammas):
        for i in range(N):
            means[i] += np.dot(g[:, i], o)
        wsum += np.sum(g, axis=0)
    means /= wsum
    sig, wsum = np.zeros(N), np.zeros(N)
    for o, g in zip(observations, gammas):
        for i in range(N):
            sig[i] += np.dot(g[:, i], (o - means[i]) ** 2)
        wsum += np.sum(g, axis=0)
    sig = np.sqrt(sig / wsum)
    if np.any(sig < np.finfo(sig.dtype).eps):
        raise RuntimeError('at least one sigma is too small to continue.')
    return means, sig


def mstep_discrete(Bnum):
    """DiscreteOutputModel.estimate normalisation (discrete.py:214-215)."""
    return Bnum / np.sum(Bnum, axis=1)[:, None]


def mstep_transition_nonrev(Cm, gamma0):
    """Non-reversible transition-matrix MLE and initial distribution
    (maximum_likelihood.py:307-320 with estimate_P -> _tmatrix_disconnected.py:110-115 ->
    msmtools transition_matrix(C, reversible=False) = C / rowsum; pi = gamma0/sum)."""
    A = Cm / Cm.sum(axis=1)[:, None]
    pi = gamma0 / gamma0.sum()
    return A, pi


def em_gaussian(oracle, observations, A, pi, means, sigmas, niter, ignore_outliers=True):
    """niter Baum-Welch iterations (fit loop, maximum_likelihood.py:379-414) with the non-reversible
    M-step.  Returns the log-likelihood history and the final parameters."""
    hist = []
    for _ in range(niter):
        st = oracle.estep_gaussian(observations, A, pi, means, sigmas, ignore_outliers)
        hist.append(st['loglik'])
        A, pi = mstep_transition_nonrev(st['C'], st['gamma0'])
        means, sigmas = mstep_gaussian(observations, st['gammas'])
    return np.array(hist), A, pi, means, sigmas
