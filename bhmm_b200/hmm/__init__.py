from .generic_hmm import HMM  # noqa: F401
