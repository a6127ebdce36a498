"""Module-level configuration, mirroring bhmm/util/config.py:24-30.

``kernel`` selects the implementation the estimators push into ``hidden.set_implementation`` and
``OutputModel.set_implementation`` (maximum_likelihood.py:142-143); here the only kernel is 'cuda'.
"""
import numpy as np

kernel = 'cuda'
dtype = np.float64
verbose = False
