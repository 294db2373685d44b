"""ctypes binding of libtopopt_b200.so (the C-ABI declared in include/topopt_b200.h).

The library is the product; this module only loads it and declares signatures.  There is no
Python/CPU fallback: if the shared library is missing or no CUDA device is present the calls fail
loudly (TobError).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TOB_LIB_PATH: load another build of the same library (A/B kernel experiments); default = the in-tree build
LIB_PATH = os.environ.get("TOB_LIB_PATH") or os.path.join(_HERE, "libtopopt_b200.so")


class TobError(RuntimeError):
    pass


class GridDesc(C.Structure):
    _fields_ = [("nxn", C.c_int), ("nyn", C.c_int), ("nzn", C.c_int), ("xc", C.c_double * 6),
                ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int),
                ("nccl_id", C.c_void_p), ("stream", C.c_void_p)]


class SolverOpts(C.Structure):
    _fields_ = [("pc", C.c_int), ("nlvls", C.c_int), ("rtol", C.c_double), ("atol", C.c_double),
                ("dtol", C.c_double), ("maxits", C.c_int), ("check_every", C.c_int), ("cheb_degree", C.c_int),
                ("coarse_its", C.c_int), ("spmv_kernel", C.c_int), ("verbose", C.c_int)]


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("rel_residual", C.c_double), ("seconds", C.c_double),
                ("spmv_calls", C.c_int), ("kernel_launches", C.c_int), ("spmv_seconds", C.c_double),
                ("spmv_timed", C.c_int)]


_lib = None

# name -> (restype, argtypes).  Every symbol include/topopt_b200.h declares must appear here
# (tests/test_abi.py cross-checks this table against the header).
_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_LL = C.c_longlong
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)
_PLL = C.POINTER(C.c_longlong)
_PP = C.POINTER(C.c_void_p)

SIGNATURES = {
    "tob_last_error": (C.c_char_p, []),
    "tob_version": (_I, []),
    "tob_ctx_create": (_I, [C.POINTER(GridDesc), _PP]),
    "tob_ctx_destroy": (_I, [_P]),
    "tob_ctx_sizes": (_I, [_P, _PLL, _PLL, _PLL, _PLL]),
    "tob_ctx_slab": (_I, [_P, _PI, _PI, _PI, _PI]),
    "tob_slab_partition": (_I, [_I, _I, _I, _PI, _PI, _PI, _PI]),
    "tob_ctx_sync": (_I, [_P]),
    "tob_ctx_launch_count": (_LL, [_P, _I]),
    "tob_ctx_stream": (_P, [_P]),
    "tob_dev_alloc": (_I, [_P, _LL, _PP]),
    "tob_dev_free": (_I, [_P, _P]),
    "tob_memcpy": (_I, [_P, _P, _P, _LL]),
    "tob_nccl_unique_id": (_I, [_P]),
    "tob_vec_scale": (_I, [_P, _P, _D, _LL]),
    "tob_elast_set_profile": (_I, [_P, _I]),
    "tob_solver_opts_default": (None, [C.POINTER(SolverOpts)]),
    "tob_elast_create": (_I, [_P, _D, C.POINTER(SolverOpts), _PP]),
    "tob_elast_destroy": (_I, [_P]),
    "tob_elast_set_bc_load": (_I, [_P, _P, _P]),
    "tob_elast_get_ke": (_I, [_P, _P]),
    "tob_elast_set_density": (_I, [_P, _P, _D, _D, _D]),
    "tob_elast_spmv": (_I, [_P, _P, _P]),
    "tob_elast_diag": (_I, [_P, _P]),
    "tob_elast_spmv_bench": (_I, [_P, _I, _I, _PD]),
    "tob_elast_solve": (_I, [_P, _P, _D, _D, _D]),
    "tob_elast_solve_obj_sens": (_I, [_P, _P, _D, _D, _D, _D, _PD, _PD, _P, _P]),
    "tob_elast_obj": (_I, [_P, _P, _D, _D, _D, _D, _PD, _PD]),
    "tob_elast_sens": (_I, [_P, _P, _D, _D, _D, _D, _P, _P]),
    "tob_elast_get_state": (_I, [_P, _P]),
    "tob_elast_set_state": (_I, [_P, _P]),
    "tob_elast_get_rhs": (_I, [_P, _P]),
    "tob_elast_last_stats": (_I, [_P, C.POINTER(SolveStats)]),
    "tob_elast_set_opts": (_I, [_P, C.POINTER(SolverOpts)]),
    "tob_filter_create": (_I, [_P, _I, _D, _PP]),
    "tob_filter_destroy": (_I, [_P]),
    "tob_filter_elemconn": (_I, [_P, _PI]),
    "tob_filter_project": (_I, [_P, _P, _P, _P, _I, _D, _D]),
    "tob_filter_gradients": (_I, [_P, _P, _P, _P, _I, _PP, _I, _D, _D]),
    "tob_filter_mnd": (_I, [_P, _P, _PD]),
    "tob_filter_increase_beta": (_I, [_PD, _D, _D, _I, _D, _PI]),
    "tob_filter_get_hs": (_I, [_P, _P]),
    "tob_filter_pde_apply": (_I, [_P, _P, _P]),
    "tob_filter_last_stats": (_I, [_P, C.POINTER(SolveStats)]),
    "tob_mma_create": (_I, [_P, _I, _P, _P, _P, _PP]),
    "tob_mma_destroy": (_I, [_P]),
    "tob_mma_set_outer_movelimit": (_I, [_P, _D, _D, _D, _P, _P, _P]),
    "tob_mma_update": (_I, [_P, _P, _P, _P, _PP, _P, _P]),
    "tob_mma_design_change": (_I, [_P, _P, _P, _PD]),
    "tob_mma_get_restart": (_I, [_P, _P, _P, _P, _P, _PI]),
    "tob_mma_set_restart": (_I, [_P, _P, _P, _P, _P, _I]),
    "tob_elast_spmv_tf32": (_I, [_P, _P, _P]),
    "tob_elast_spmv_tf32_bench": (_I, [_P, _I, _PD]),
    "tob_elast_gmg_nlevels": (_I, [_P, _PI]),
    "tob_elast_gmg_level_stencil": (_I, [_P, _I, _P, _PI, _PI]),
    "tob_microbench_fp64": (_I, [_P, _I, _PD]),
    "tob_microbench_copy": (_I, [_P, _LL, _PD]),
}


def load():
    """Load the shared library (once) and declare the C-ABI signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TobError(f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                       "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().tob_last_error()
        raise TobError(msg.decode() if msg else f"libtopopt_b200 call failed with code {rc}")
