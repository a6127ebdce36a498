"""bhmm_b200.hidden -- the bhmm.hidden surface (bhmm/hidden/__init__.py:22) on sm_100a CUDA kernels."""
from .api import *  # noqa: F401,F403
from .api import __all__  # noqa: F401
