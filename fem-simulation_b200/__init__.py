"""femb200: B200-native implicit / quasi-static hexahedral FEM multigrid step.

Drop-in for the hot path of caixiao-0725/fem-simulation: the Python
simulation API (``Object``) on top of hand-written sm_100a CUDA kernels behind the C ABI of
``include/femb200.h`` (``libfemb200.so``).  There is no CPU fallback: constructing an
``Object`` without the CUDA library and a GPU raises.
"""
from . import _lib, hierarchy, mesh  # noqa: F401
from .object import Object  # noqa: F401
from .object_dynamic import DynamicObject  # noqa: F401
from .cloth import Cloth  # noqa: F401

__all__ = ["Object", "DynamicObject", "Cloth", "mesh", "hierarchy", "_lib"]
