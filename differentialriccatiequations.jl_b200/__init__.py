"""B200-native LRSIF-ADI hot path of DifferentialRiccatiEquations.jl.

Import as ``import dre_b200`` (shim at the repo root).  Submodules:
  pencils   synthetic Rail-shaped / 3D-heat input data
  capi      ctypes binding of the C ABI (include/dre_b200.h, libdre_b200.so)
  api       host-side mirror of the reference's Julia API over the C ABI
"""
from . import pencils  # noqa: F401
from . import capi  # noqa: F401
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
