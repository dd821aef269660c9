"""multioutputihgp_b200 - B200-native (sm_100a) hot path of lim271/MultiOutputIHGP.

``MOIHGP`` and ``MOIHGPOnlineLearning`` mirror the reference's Python entry points
(moihgp/__init__.py:1-6); ``MOIHGPSequences`` is the whole-sequence interface to the CUDA
kernels.  Importing the package does not need a GPU; creating a model does (no CPU fallback).
"""
from .pywrapper import MOIHGP
from .online_learning import MOIHGPOnlineLearning
from .batched import MOIHGPSequences, SMOOTH_NONE, SMOOTH_REFERENCE_LITERAL, SMOOTH_RTS

__all__ = ["MOIHGP", "MOIHGPOnlineLearning", "MOIHGPSequences", "SMOOTH_NONE", "SMOOTH_REFERENCE_LITERAL", "SMOOTH_RTS"]
