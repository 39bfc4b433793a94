"""B200-native (sm_100a) implementation of the hot path of GBATZOLIS/conditional_score_diffusion:
the NCSN++ score network and the reverse-SDE predictor-corrector sampler, behind the reference's
own Python module surface (models.ncsnpp.NCSNpp, sde_lib, sampling.*, losses, op.upfirdn2d).

All arithmetic runs in hand-written CUDA kernels reached through the C ABI in include/csd_b200.h
(libcsd_b200.so, built in-tree by `conditional_score_diffusion_b200.build`).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
