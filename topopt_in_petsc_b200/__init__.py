"""topopt_in_petsc_b200 -- B200-native inner loop for TopOpt_in_PETSc's hot path.

The product is libtopopt_b200.so (hand-written sm_100a CUDA behind the C-ABI in
include/topopt_b200.h) plus the C++ host classes under host/.  This Python package is only the
ctypes mirror used by tests and bench.py.
"""
from ._lib import TobError, load, LIB_PATH  # noqa: F401
from .api import Context, DevVec, Filter, LinearElasticity, MMA, nccl_unique_id, solver_opts  # noqa: F401
