"""Batched, device-resident driver of the fused kernels (include/bhmm_b200.h, group 3).

``TrajectoryBatch`` keeps every observation trajectory of a data set concatenated in GPU memory and runs

* one Baum-Welch E-step over all of them (emission + scaled forward + backward + gamma/xi statistics; what
  MaximumLikelihoodEstimator._forward_backward does per trajectory, bhmm/estimators/maximum_likelihood.py:221-282),
* Viterbi paths of all trajectories (compute_viterbi_paths, :332-352),
* one Gibbs hidden-path sweep (BayesianHMMSampler._updateHiddenStateTrajectories,
  bhmm/estimators/bayesian_sampling.py:283-331) with the path statistics of bhmm/hmm/generic_hmm.py:297-334,398-431

per call.  PyTorch is used only to own device buffers and the stream (and NCCL in ``bhmm_b200.dist``); all
arithmetic happens in libbhmm_b200.so.  There is no CPU path.
"""
import ctypes as C

import os

import numpy as np

from . import _lib
from ._lib import lib, check, dptr, f64


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.CudaUnavailableError('bhmm_b200.engine needs a CUDA device (no CPU fallback)')
    return torch


def unpack_stats(stats, N):
    """Split the packed E-step statistics [loglik | gamma0 | C | sum gamma | sum gamma d | sum gamma d^2]."""
    s = np.asarray(stats, dtype=np.float64)
    o = 0
    out = {'loglik': float(s[0])}
    o = 1
    out['gamma0'] = s[o:o + N].copy(); o += N
    out['C'] = s[o:o + N * N].reshape(N, N).copy(); o += N * N
    out['wsum'] = s[o:o + N].copy(); o += N
    out['wd'] = s[o:o + N].copy(); o += N
    out['wdd'] = s[o:o + N].copy()
    return out


class TrajectoryBatch(object):
    """All trajectories of one data set (or of one rank's shard of it), resident on one GPU.

    Parameters
    ----------
    observations : list of 1-D arrays
        float arrays for a Gaussian output model, integer arrays for a discrete one.
    nstates : int
    device : torch device or None (current CUDA device)
    chunk, warm : int
        frames per chain / warm-up frames of the time-chunked kernels; 0 = automatic.
    """

    def __init__(self, observations, nstates, device=None, chunk=0, warm=0):
        first = np.asarray(observations[0])
        host_dtype = np.int32 if np.issubdtype(first.dtype, np.integer) else np.float64
        lengths = [len(o) for o in observations]
        if len(lengths) == 0 or min(lengths) <= 0:
            raise ValueError('every trajectory needs at least one frame')
        cat = np.concatenate([np.asarray(o, dtype=host_dtype) for o in observations])
        self._setup(cat, lengths, nstates, device, chunk, warm)

    @classmethod
    def from_concatenated(cls, rows, lengths, nstates, device=None, chunk=0, warm=0):
        """Build from one concatenated per-frame array (numpy, or a torch tensor that may already live on the GPU)
        and the list of trajectory lengths."""
        self = cls.__new__(cls)
        self._setup(rows, lengths, nstates, device, chunk, warm)
        return self

    def _setup(self, cat, lengths, nstates, device, chunk, warm):
        torch = _torch()
        self.torch = torch
        self.N = int(nstates)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        lengths = np.asarray(lengths, dtype=np.int64)
        if len(lengths) == 0 or np.any(lengths <= 0):
            raise ValueError('every trajectory needs at least one frame')
        self.K = len(lengths)
        self.lengths = lengths
        self.offsets = np.zeros(self.K + 1, dtype=np.int64)
        np.cumsum(lengths, out=self.offsets[1:])
        self.rows = int(self.offsets[-1])
        if not torch.is_tensor(cat):
            cat = torch.from_numpy(np.ascontiguousarray(cat))
        if cat.numel() != self.rows:
            raise ValueError('concatenated observations have %d frames, lengths sum to %d' % (cat.numel(), self.rows))
        self.discrete = not cat.dtype.is_floating_point
        with torch.cuda.device(self.device):
            self.obs = cat.to(device=self.device, dtype=torch.int32 if self.discrete else torch.float64).contiguous().clone() \
                if cat.is_cuda else cat.to(device=self.device, dtype=torch.int32 if self.discrete else torch.float64)
            self._handle = C.c_void_p()
            rc = lib.bhmm_b200_batch_create(C.byref(self._handle), self.offsets.ctypes.data_as(C.POINTER(C.c_longlong)),
                                            self.K, self.N, int(chunk), int(warm))
            check(rc)
            self._attach()
        self._stats = torch.zeros(lib.bhmm_b200_stats_len_gaussian(self.N), dtype=torch.float64, device=self.device)
        self._path = None
        self._counts = None
        self._sums = None

    # -------------------------------------------------------------------------------------------- plumbing
    def _attach(self):
        nbytes = int(lib.bhmm_b200_batch_workspace_bytes(self._handle))
        self.workspace_bytes = nbytes
        if os.environ.get('BHMM_B200_OWN_WORKSPACE'):
            # debugging aid (compute-sanitizer initcheck needs a fresh cudaMalloc): the library allocates its own arena
            self._workspace = None
            return
        self._workspace = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)
        check(lib.bhmm_b200_batch_attach_workspace(self._handle, C.c_void_p(self._workspace.data_ptr()), nbytes))
        self.workspace_bytes = nbytes

    def replan(self, chunk=0, warm=0):
        """Change the time-chunking (frames per chain, warm-up frames) and re-carve the workspace."""
        check(lib.bhmm_b200_batch_replan(self._handle, int(chunk), int(warm)))
        self._attach()

    def set_observations(self, host_array, non_blocking=True):
        """Host -> device copy of a new concatenated observation array of the same shape (end-to-end benchmarks)."""
        self.obs.copy_(host_array, non_blocking=non_blocking)

    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle.value:
            lib.bhmm_b200_batch_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        info = np.zeros(8)
        lib.bhmm_b200_batch_info(self._handle, dptr(info))
        return dict(chains=int(info[0]), chunk=int(info[1]), warm=int(info[2]), fixups_fwd=int(info[3]),
                    fixups_bwd=int(info[4]), worst_fwd=float(info[5]), worst_bwd=float(info[6]), rerun=int(info[7]))

    @property
    def uses_lane_kernels(self):
        """True when the small-N one-thread-per-chain kernels run (N <= 16), False for the general-N team kernels."""
        return bool(lib.bhmm_b200_batch_uses_lane_kernels(self._handle))

    def set_profiling(self, on=True):
        """Record CUDA events around the forward and the backward+statistics kernels of every E-step."""
        check(lib.bhmm_b200_batch_set_profiling(self._handle, int(bool(on))))

    def kernel_ms(self):
        """Device times of the last E-step: dict(forward, backward_stats, span) in milliseconds."""
        ms = np.zeros(4)
        lib.bhmm_b200_batch_kernel_ms(self._handle, dptr(ms))
        return dict(forward=float(ms[0]), backward_stats=float(ms[1]), span=float(ms[2]))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def split(self, flat):
        """Cut a concatenated per-frame array back into the list of per-trajectory arrays."""
        return [flat[self.offsets[k]:self.offsets[k + 1]] for k in range(self.K)]

    # -------------------------------------------------------------------------------------------- E-step
    def estep_gaussian(self, A, pi, means, sigmas, ignore_outliers=True, gamma_out=None):
        """One E-step; returns the packed statistics as a DEVICE tensor (see ``unpack_stats``).

        ``gamma_out``: optional (rows,N) float64 CUDA tensor that receives the state probabilities.
        """
        A_, pi_, m_, s_ = f64(A), f64(pi), f64(means), f64(sigmas)
        g = C.c_void_p(gamma_out.data_ptr()) if gamma_out is not None else None
        with self.torch.cuda.device(self.device):
            rc = lib.bhmm_b200_estep_gaussian(self._handle, C.c_void_p(self.obs.data_ptr()), dptr(A_), dptr(pi_),
                                              dptr(m_), dptr(s_), int(bool(ignore_outliers)), g,
                                              C.c_void_p(self._stats.data_ptr()), self._stream())
        check(rc)
        return self._stats

    def estep_discrete(self, A, pi, B, ignore_outliers=False, gamma_out=None):
        """One E-step for a discrete output model; returns (stats, Bnum) device tensors."""
        A_, pi_, B_ = f64(A), f64(pi), f64(B)
        N, M = B_.shape
        if getattr(self, '_Bnum', None) is None or tuple(self._Bnum.shape) != (N, M):
            self._Bnum = self.torch.zeros((N, M), dtype=self.torch.float64, device=self.device)
        g = C.c_void_p(gamma_out.data_ptr()) if gamma_out is not None else None
        with self.torch.cuda.device(self.device):
            rc = lib.bhmm_b200_estep_discrete(self._handle, C.c_void_p(self.obs.data_ptr()), dptr(A_), dptr(pi_),
                                              dptr(B_), M, int(bool(ignore_outliers)), g,
                                              C.c_void_p(self._stats.data_ptr()), C.c_void_p(self._Bnum.data_ptr()),
                                              self._stream())
        check(rc)
        return self._stats, self._Bnum

    # -------------------------------------------------------------------------------------------- Viterbi
    def _path_buffer(self):
        if self._path is None:
            self._path = self.torch.zeros(self.rows, dtype=self.torch.int32, device=self.device)
        return self._path

    def viterbi_gaussian(self, A, pi, means, sigmas, ignore_outliers=True):
        """Viterbi paths of all trajectories as one concatenated int32 DEVICE tensor."""
        A_, pi_, m_, s_ = f64(A), f64(pi), f64(means), f64(sigmas)
        path = self._path_buffer()
        with self.torch.cuda.device(self.device):
            rc = lib.bhmm_b200_viterbi_gaussian(self._handle, C.c_void_p(self.obs.data_ptr()), dptr(A_), dptr(pi_),
                                                dptr(m_), dptr(s_), int(bool(ignore_outliers)),
                                                C.c_void_p(path.data_ptr()), self._stream())
        check(rc)
        return path

    def viterbi_discrete(self, A, pi, B, ignore_outliers=False):
        A_, pi_, B_ = f64(A), f64(pi), f64(B)
        path = self._path_buffer()
        with self.torch.cuda.device(self.device):
            rc = lib.bhmm_b200_viterbi_discrete(self._handle, C.c_void_p(self.obs.data_ptr()), dptr(A_), dptr(pi_),
                                                dptr(B_), B_.shape[1], int(bool(ignore_outliers)),
                                                C.c_void_p(path.data_ptr()), self._stream())
        check(rc)
        return path

    # -------------------------------------------------------------------------------------------- Gibbs
    def _gibbs_buffers(self):
        N = self.N
        if self._counts is None:
            self._counts = self.torch.zeros(N * N + 2 * N, dtype=self.torch.int64, device=self.device)
            self._sums = self.torch.zeros(2 * N, dtype=self.torch.float64, device=self.device)
        return self._counts, self._sums

    def gibbs_gaussian(self, A, pi, means, sigmas, seed=0, sweep=0, uniforms=None, ignore_outliers=True):
        """One hidden-path sweep.  Returns (path, counts, sums, loglik): concatenated int32 paths, int64
        [C (N*N) | n0 (N) | frames per state (N)], float64 [sum o (N) | sum o^2 (N)] (all DEVICE tensors) and the
        log-likelihood of the forward pass.  ``uniforms``: optional (rows,) float64 CUDA tensor, one draw per
        frame (parity with the reference's glibc stream); default is device Philox keyed by (seed, sweep)."""
        A_, pi_, m_, s_ = f64(A), f64(pi), f64(means), f64(sigmas)
        path = self._path_buffer()
        counts, sums = self._gibbs_buffers()
        ll = C.c_double(0.0)
        u = C.c_void_p(uniforms.data_ptr()) if uniforms is not None else None
        with self.torch.cuda.device(self.device):
            rc = lib.bhmm_b200_gibbs_gaussian(self._handle, C.c_void_p(self.obs.data_ptr()), dptr(A_), dptr(pi_),
                                              dptr(m_), dptr(s_), int(bool(ignore_outliers)), u,
                                              C.c_ulonglong(int(seed)), C.c_ulonglong(int(sweep)),
                                              C.c_void_p(path.data_ptr()), C.c_void_p(counts.data_ptr()),
                                              C.c_void_p(sums.data_ptr()), C.byref(ll), self._stream())
        check(rc)
        return path, counts, sums, ll.value

    def gibbs_discrete(self, A, pi, B, seed=0, sweep=0, uniforms=None, ignore_outliers=False):
        """Discrete counterpart; returns (path, counts, symbol histogram (N,M) int64, loglik)."""
        A_, pi_, B_ = f64(A), f64(pi), f64(B)
        N, M = B_.shape
        path = self._path_buffer()
        counts, _ = self._gibbs_buffers()
        if getattr(self, '_hist', None) is None or tuple(self._hist.shape) != (N, M):
            self._hist = self.torch.zeros((N, M), dtype=self.torch.int64, device=self.device)
        ll = C.c_double(0.0)
        u = C.c_void_p(uniforms.data_ptr()) if uniforms is not None else None
        with self.torch.cuda.device(self.device):
            rc = lib.bhmm_b200_gibbs_discrete(self._handle, C.c_void_p(self.obs.data_ptr()), dptr(A_), dptr(pi_),
                                              dptr(B_), M, int(bool(ignore_outliers)), u, C.c_ulonglong(int(seed)),
                                              C.c_ulonglong(int(sweep)), C.c_void_p(path.data_ptr()),
                                              C.c_void_p(counts.data_ptr()), C.byref(ll), self._stream())
            check(rc)
            self._hist.zero_()
            rc = lib.bhmm_b200_path_symbol_histogram(C.c_void_p(path.data_ptr()), C.c_void_p(self.obs.data_ptr()),
                                                     C.c_longlong(self.rows), N, M,
                                                     C.c_void_p(self._hist.data_ptr()), self._stream())
        check(rc)
        return path, counts, self._hist, ll.value

    def unpack_counts(self, counts):
        c = counts.cpu().numpy()
        N = self.N
        return dict(C=c[:N * N].reshape(N, N).copy(), n0=c[N * N:N * N + N].copy(), count=c[N * N + N:].copy())
