"""Estimators with the interface of bhmm/estimators, driving the batched CUDA engine."""
from .maximum_likelihood import MaximumLikelihoodEstimator  # noqa: F401
from .bayesian_sampling import BayesianHMMSampler  # noqa: F401
