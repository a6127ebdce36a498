"""Base class of the emission models; mirrors bhmm/output_models/outputmodel.py:24-150."""
import warnings

import numpy as np


class OutputModel(object):
    """HMM output probability model base class (bhmm/output_models/outputmodel.py:24-66).

    Parameters
    ----------
    nstates : int
        number of hidden states
    ignore_outliers : bool
        if True, observations that have zero probability under every state get probability 1 for every state
        (outputmodel.py:119-131); ``found_outliers`` records that this happened.
    """
    __IMPL_CUDA__ = 2
    __impl__ = __IMPL_CUDA__

    def __init__(self, nstates, ignore_outliers=True):
        self._nstates = nstates
        self.ignore_outliers = ignore_outliers
        self.found_outliers = False

    @property
    def nstates(self):
        r""" Number of hidden states """
        return self._nstates

    def set_implementation(self, impl):
        """'cuda' only (outputmodel.py:69-86 knows 'python' and 'c'); other names warn and keep 'cuda'."""
        if impl.lower() != 'cuda':
            warnings.warn('Implementation ' + impl + ' is not provided by bhmm_b200. Using the cuda implementation.')
        self.__impl__ = self.__IMPL_CUDA__

    def log_p_obs(self, obs, out=None, dtype=np.float32):
        """Element-wise logarithm of p_obs (outputmodel.py:93-117)."""
        if out is None:
            return np.log(self.p_obs(obs))
        self.p_obs(obs, out=out)
        np.log(out, out=out)
        return out
