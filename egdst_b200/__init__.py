"""egdst_b200 -- B200-native solver and simulator hot paths for egdst models (DC-EGM).

Host side: ``EgdstModel`` mirrors the reference's ``@egdstmodel`` class; ``codegen`` restates
``compile.m``; ``build`` drives nvcc (sm_100a); ``capi`` binds the C-ABI of include/egdst_b200.h.
The compute lives in ``csrc/`` (hand-written CUDA).  There is no CPU fallback.
"""
from .model import EgdstModel  # noqa: F401
from . import examples  # noqa: F401

__all__ = ["EgdstModel", "examples"]
