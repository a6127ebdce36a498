"""Register the 'cuda' implementation inside the reference package (when `bhmm` is importable).

After ``bhmm_b200.install()``:
    bhmm.hidden.set_implementation('cuda')          (or bhmm.config.kernel = 'cuda' before building an estimator)
routes bhmm.hidden.{forward, backward, state_probabilities, state_counts, transition_counts, viterbi,
sample_path} to the CUDA kernels, and ``OutputModel.set_implementation('cuda')`` routes p_obs to them, so the
reference's MaximumLikelihoodEstimator and BayesianHMMSampler run unchanged on top.
INTEGRATION.md shows the equivalent source patch for bhmm/hidden/api.py.
"""
import functools

from . import hidden as cuda_hidden

_IMPL_CUDA = 2


def install(bhmm_module=None, device_cache=True):
    """Patch ``bhmm.hidden.api`` / ``bhmm.hidden`` / ``bhmm.output_models`` in place; returns the bhmm module.

    ``device_cache``: switch on the buffer-identity device cache of ``bhmm_b200.hidden``: the
    reference estimators hand the same private preallocated arrays from call to call, so the tables this library wrote
    into them need not be uploaded again (11 -> 4 (T,N) PCIe transfers per trajectory and EM iteration)."""
    cuda_hidden.set_device_cache(device_cache)
    if bhmm_module is None:
        import bhmm as bhmm_module  # the reference package must be importable
    import importlib
    api = importlib.import_module(bhmm_module.__name__ + '.hidden.api')
    hidden_pkg = importlib.import_module(bhmm_module.__name__ + '.hidden')
    if getattr(api, '__bhmm_b200_installed__', False):
        return bhmm_module
    api.__IMPL_CUDA__ = _IMPL_CUDA
    orig_set = api.set_implementation

    def set_implementation(impl):
        if impl.lower() == 'cuda':
            api.__impl__ = _IMPL_CUDA
        else:
            orig_set(impl)
    set_implementation.__doc__ = orig_set.__doc__
    api.set_implementation = set_implementation
    hidden_pkg.set_implementation = set_implementation

    for name in ('forward', 'backward', 'state_probabilities', 'state_counts', 'transition_counts', 'viterbi',
                 'sample_path'):
        orig = getattr(api, name)
        ours = getattr(cuda_hidden, name)

        def make(orig, ours):
            @functools.wraps(orig)
            def dispatch(*args, **kwargs):
                if api.__impl__ == _IMPL_CUDA:
                    return ours(*args, **kwargs)
                return orig(*args, **kwargs)
            return dispatch
        wrapped = make(orig, ours)
        setattr(api, name, wrapped)
        setattr(hidden_pkg, name, wrapped)

    # emission side: OutputModel.set_implementation only knows 'python' / 'c' (outputmodel.py:69-86)
    om = importlib.import_module(bhmm_module.__name__ + '.output_models.outputmodel')
    gm = importlib.import_module(bhmm_module.__name__ + '.output_models.gaussian')
    dm = importlib.import_module(bhmm_module.__name__ + '.output_models.discrete')
    from .output_models import GaussianOutputModel as CudaGaussian, DiscreteOutputModel as CudaDiscrete
    base_set = om.OutputModel.set_implementation

    def om_set_implementation(self, impl):
        if impl.lower() == 'cuda':
            self.__impl__ = _IMPL_CUDA
        else:
            base_set(self, impl)
    om.OutputModel.set_implementation = om_set_implementation

    g_p_obs = gm.GaussianOutputModel.p_obs

    def gaussian_p_obs(self, obs, out=None):
        if getattr(self, '__impl__', None) == _IMPL_CUDA:
            tmp = CudaGaussian(self.nstates, means=self.means, sigmas=self.sigmas, ignore_outliers=self.ignore_outliers)
            res = tmp.p_obs(obs, out=out)
            self.found_outliers = self.found_outliers or tmp.found_outliers
            return res
        return g_p_obs(self, obs, out=out)
    gm.GaussianOutputModel.p_obs = gaussian_p_obs

    d_p_obs = dm.DiscreteOutputModel.p_obs

    def discrete_p_obs(self, obs, out=None):
        if getattr(self, '__impl__', None) == _IMPL_CUDA:
            tmp = CudaDiscrete(self.output_probabilities, ignore_outliers=self.ignore_outliers)
            return tmp.p_obs(obs, out=out)
        return d_p_obs(self, obs, out=out)
    dm.DiscreteOutputModel.p_obs = discrete_p_obs

    # DiscreteOutputModel.estimate dispatches its scatter-add on the implementation too (discrete.py:205-213 ->
    # impl_c/_discrete.c:1-32 _update_pout): 'cuda' runs the drop-in bhmm_b200_discrete_update_pout per trajectory
    d_estimate = dm.DiscreteOutputModel.estimate

    def discrete_estimate(self, observations, weights):
        if getattr(self, '__impl__', None) != _IMPL_CUDA:
            return d_estimate(self, observations, weights)
        import numpy as np
        from ._lib import lib, dptr, iptr, check
        N, M = self._output_probabilities.shape
        B = np.zeros((N, M))
        for k in range(len(observations)):
            sym = np.ascontiguousarray(observations[k], dtype=np.int32)
            w = np.ascontiguousarray(weights[k], dtype=np.float64)
            lib.bhmm_b200_discrete_update_pout(iptr(sym), dptr(w), sym.shape[0], N, M, dptr(B))
            check()
        B /= np.sum(B, axis=1)[:, None]
        self._output_probabilities = B
    dm.DiscreteOutputModel.estimate = discrete_estimate

    api.__bhmm_b200_installed__ = True
    return bhmm_module
