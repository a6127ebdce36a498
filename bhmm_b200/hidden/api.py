"""Drop-in counterpart of bhmm/hidden/api.py with a single implementation: 'cuda'.

Same function names, argument meaning, array layouts (C-contiguous float64, (T,N) time-major, int32 paths),
in-place ``*_out`` semantics and exceptions as the reference dispatcher (bhmm/hidden/api.py:44-304) and its
Cython wrappers (bhmm/hidden/impl_c/hidden.pyx:39-204).  Every call copies its host arrays to the GPU, runs the
hand-written kernels behind libbhmm_b200.so and returns completed numpy results.  There is no CPU path: without
a usable CUDA device the calls raise ``CudaUnavailableError``.

For data that should stay on the GPU across calls use ``bhmm_b200.engine.TrajectoryBatch`` instead (one fused
E-step / Viterbi / Gibbs sweep over all trajectories).
"""
import ctypes
import warnings

import numpy as np

from .. import _lib
from .._lib import lib, dptr, iptr, f64, check
from ..util import config

__all__ = ['set_implementation', 'forward', 'backward', 'state_probabilities', 'state_counts',
           'transition_counts', 'viterbi', 'sample_path']

# implementation codes (bhmm/hidden/api.py:36-41 has python=0, c=1; 'cuda' is the one added here)
__IMPL_CUDA__ = 2
__impl__ = __IMPL_CUDA__


def set_implementation(impl):
    """Sets the implementation of this module (bhmm/hidden/api.py:44-62).

    Parameters
    ----------
    impl : str
        'cuda' (case-insensitive).  Like the reference, an unknown name only warns; but the implementation that
        stays selected is the CUDA one -- this package ships no 'python' or 'c' kernels and never falls back to
        the CPU.
    """
    global __impl__
    if impl.lower() != 'cuda':
        warnings.warn('Implementation ' + impl + ' is not provided by bhmm_b200. Using the cuda implementation.')
    __impl__ = __IMPL_CUDA__


def _require_f64():
    if config.dtype != np.float64:
        raise TypeError('bhmm_b200 computes in float64 only (config.dtype = %r)' % (config.dtype,))


def _out_buffer(out, T, N):
    """Returns (array to hand to C, array to return).  In-place when `out` is C-contiguous float64."""
    if out is None:
        buf = np.zeros((T, N), dtype=np.float64, order='C')
        return buf, buf
    if isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags['C_CONTIGUOUS'] and out.flags['WRITEABLE']:
        return out, out
    buf = np.zeros((T, N), dtype=np.float64, order='C')
    return buf, out


def forward(A, pobs, pi, T=None, alpha_out=None):
    """Compute P( obs | A, B, pi ) and all forward coefficients (bhmm/hidden/api.py:65-96, _hidden.c:16-66).

    Returns ``(logprob, alpha)``; ``alpha_out`` (at least T rows) is filled in place and returned when given.
    """
    _require_f64()
    if T is None:
        T = pobs.shape[0]
    elif T > pobs.shape[0]:
        raise TypeError('T must be at most the length of pobs.')
    N = A.shape[0]
    if alpha_out is not None and T > alpha_out.shape[0]:
        raise TypeError('alpha_out must at least have length T in order to fit trajectory.')
    buf, ret = _out_buffer(alpha_out, T, N)
    A_, pobs_, pi_ = f64(A), f64(pobs), f64(pi)
    logprob = lib.bhmm_b200_forward(dptr(buf), dptr(A_), dptr(pobs_), dptr(pi_), N, T)
    check()
    if buf is not ret:
        ret[:T] = buf[:T]
    return logprob, ret


def backward(A, pobs, T=None, beta_out=None):
    """Compute all backward coefficients, with scaling (bhmm/hidden/api.py:99-125, _hidden.c:69-110)."""
    _require_f64()
    if T is None:
        T = pobs.shape[0]
    elif T > pobs.shape[0]:
        raise ValueError('T must be at most the length of pobs.')
    N = A.shape[0]
    if beta_out is not None and T > beta_out.shape[0]:
        raise ValueError('beta_out must at least have length T in order to fit trajectory.')
    buf, ret = _out_buffer(beta_out, T, N)
    A_, pobs_ = f64(A), f64(pobs)
    lib.bhmm_b200_backward(dptr(buf), dptr(A_), dptr(pobs_), N, T)
    check()
    if buf is not ret:
        ret[:T] = buf[:T]
    return ret


def state_probabilities(alpha, beta, T=None, gamma_out=None):
    """(T,N) matrix of the probabilities of being in state i at time t (bhmm/hidden/api.py:133-188)."""
    if alpha.shape[0] != beta.shape[0]:
        raise ValueError('Inconsistent sizes of alpha and beta.')
    if T is None:
        T = alpha.shape[0] if gamma_out is None else gamma_out.shape[0]
    N = alpha.shape[1]
    if gamma_out is None:
        rows = alpha.shape[0]
    elif gamma_out.shape[0] < alpha.shape[0]:
        rows = T
        if gamma_out.shape[0] != rows:
            raise ValueError('gamma_out has %d rows, expected T = %d' % (gamma_out.shape[0], rows))
    else:
        rows = alpha.shape[0]
        if gamma_out.shape[0] != rows:
            raise ValueError('gamma_out has %d rows, expected %d' % (gamma_out.shape[0], rows))
    buf, ret = _out_buffer(gamma_out, rows, N)
    a_, b_ = f64(alpha), f64(beta)
    check(lib.bhmm_b200_state_probabilities(dptr(buf), dptr(a_), dptr(b_), N, rows))
    if gamma_out is None:
        return buf[:T] if T < rows else buf
    if buf is not ret:
        ret[:rows] = buf
    return ret


def state_counts(gamma, T, out=None):
    """Sum of the probabilities of being in state i over t < T (bhmm/hidden/api.py:191-211)."""
    g_ = f64(gamma)
    T = min(int(T), g_.shape[0])
    N = g_.shape[1]
    res = np.zeros(N, dtype=np.float64)
    check(lib.bhmm_b200_state_counts(dptr(res), dptr(g_), N, T))
    if out is not None:
        out[...] = res
        return out
    return res


def transition_counts(alpha, beta, A, pobs, T=None, out=None):
    """Sum over t of the probability to transition from i to j (bhmm/hidden/api.py:214-248, _hidden.c:148-183)."""
    _require_f64()
    if T is None:
        T = pobs.shape[0]
    elif T > pobs.shape[0]:
        raise ValueError('T must be at most the length of pobs.')
    N = len(A)
    if out is None or not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags['C_CONTIGUOUS']):
        Cbuf = np.zeros((N, N), dtype=np.float64, order='C')
    else:
        Cbuf = out
    a_, b_, A_, p_ = f64(alpha), f64(beta), f64(A), f64(pobs)
    rc = lib.bhmm_b200_transition_counts(dptr(Cbuf), dptr(A_), dptr(p_), dptr(a_), dptr(b_), N, T)
    check(rc)
    if out is not None and Cbuf is not out:
        out[...] = Cbuf
        return out
    return Cbuf


def viterbi(A, pobs, pi):
    """Maximum-likelihood hidden path (bhmm/hidden/api.py:251-274, _hidden.c:203-281); int32, bit-exact."""
    _require_f64()
    N = A.shape[0]
    T = pobs.shape[0]
    path = np.zeros(T, dtype=ctypes.c_int, order='C')
    A_, p_, pi_ = f64(A), f64(pobs), f64(pi)
    check(lib.bhmm_b200_viterbi(iptr(path), dptr(A_), dptr(p_), dptr(pi_), N, T))
    return path


def sample_path(alpha, A, pobs, T=None, seed=None):
    """Sample the hidden path from P(S | parameters, observations) (bhmm/hidden/api.py:277-304, _hidden.c:330-378).

    ``seed`` reseeds the library's restatement of glibc srand()/rand(), so a given seed reproduces the reference's
    draws bit for bit; without a seed the stream continues.
    """
    _require_f64()
    if seed is not None:
        lib.bhmm_b200_set_seed(int(seed))
    N = pobs.shape[1]
    if T is None:
        T = pobs.shape[0]
    elif T > pobs.shape[0] or T > alpha.shape[0]:
        raise ValueError('T must be at most the length of pobs and alpha.')
    path = np.zeros(T, dtype=ctypes.c_int, order='C')
    a_, A_ = f64(alpha), f64(A)
    rc = lib.bhmm_b200_sample_path(iptr(path), dptr(a_), dptr(A_), None, N, T)
    check(rc)
    return path
