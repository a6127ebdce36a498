"""Output (emission) models with the interface of bhmm/output_models (outputmodel.py:24-150)."""
from .outputmodel import OutputModel  # noqa: F401
from .gaussian import GaussianOutputModel  # noqa: F401
from .discrete import DiscreteOutputModel  # noqa: F401
