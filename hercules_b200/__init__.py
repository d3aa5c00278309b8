"""hercules_b200 -- B200-native (sm_100a) implementation of the explicit time-stepping hot path
of Hercules' quake/forward solver, behind a C ABI (include/hercules_gpu.h).

Only what the path needs lives here:
  csrc/      CUDA kernels, host-side index builders and the C ABI (libhercules_gpu.so)
  solver.py  host-side mirror of the reference's solver_* call sequence over that ABI
  _lib.py    ctypes loader (fails loudly when the library or the GPU is missing)
"""
from ._lib import build, lib, SO  # noqa: F401
from .solver import (  # noqa: F401
    Solver, HostMesh, MsgList, HerculesGpuError, PinnedArray,
    RAYLEIGH, MASS, NONE, BKT, CONVENTIONAL, EFFECTIVE, TM1, TM2, TM3, FORCE, CONV_SHEAR_1, CONV_SHEAR_2, CONV_KAPPA_1, CONV_KAPPA_2,
    FLAG_NO_FUSE, FLAG_TIMERS, FLAG_NO_OVERLAP, FLAG_TAIL_OVERLAP, FLAG_WPASS, FLAG_NO_STRUCT, FLAG_STRUCT, FLAG_DENSE_K,
)
