"""Sampling package (the reference's sampling/__init__.py is empty; the functions the Lightning
modules import live in sampling.unconditional / sampling.conditional)."""
from .conditional import get_conditional_sampling_fn, get_pc_conditional_sampler  # noqa: F401
from .correctors import get_corrector, register_corrector  # noqa: F401
from .predictors import get_predictor, register_predictor  # noqa: F401
from .unconditional import get_pc_sampler, get_sampling_fn  # noqa: F401
