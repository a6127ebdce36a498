"""bhmm_b200 -- the HMM dynamic-programming hot path of bhmm (bhmm.hidden + OutputModel.p_obs) on B200 (sm_100a).

    bhmm_b200.hidden            drop-in bhmm.hidden API ('cuda' implementation, host numpy arrays in and out)
    bhmm_b200.engine            batched device-resident E-step / Viterbi / Gibbs sweep over all trajectories
    bhmm_b200.estimators        MaximumLikelihoodEstimator / BayesianHMMSampler on top of the engine
    bhmm_b200.output_models     Gaussian / discrete emission models
    bhmm_b200.install()         registers 'cuda' inside an importable reference `bhmm` package
    bhmm_b200.estimate_hmm / bayesian_hmm / init_hmm / lag_observations   the reference's user entry points (bhmm/api.py)

Importing the package loads libbhmm_b200.so and fails loudly if it was not built; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (loads the CUDA library or raises)
from . import hidden  # noqa: F401
from .hmm import HMM  # noqa: F401
from .output_models import GaussianOutputModel, DiscreteOutputModel  # noqa: F401
from .install import install  # noqa: F401
from .api import (estimate_hmm, bayesian_hmm, init_hmm, init_gaussian_hmm, init_discrete_hmm, lag_observations,  # noqa: F401
                  gaussian_hmm, discrete_hmm)

__version__ = '0.1.0'
