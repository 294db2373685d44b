"""ctypes wrapper of oracle/_ref/libmma_ref.so = the reference's own unmodified MMA.cc (built by
oracle/Makefile from /root/reference against the stub petsc.h).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import numpy as np

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libmma_ref.so")


def available():
    return os.path.exists(SO)


class MMARef:
    def __init__(self, n, m):
        l = C.CDLL(SO)
        l.mmaref_create.restype = C.c_void_p
        l.mmaref_create.argtypes = [C.c_int, C.c_int]
        l.mmaref_destroy.argtypes = [C.c_void_p]
        l.mmaref_set_outer_movelimit.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 3
        l.mmaref_update.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        l.mmaref_design_change.restype = C.c_double
        l.mmaref_design_change.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.l, self.n, self.m = l, n, m
        self.h = l.mmaref_create(n, m)

    def set_outer_movelimit(self, Xmin, Xmax, movlim, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lo, hi = np.empty(self.n), np.empty(self.n)
        self.l.mmaref_set_outer_movelimit(self.h, Xmin, Xmax, movlim, x.ctypes.data, lo.ctypes.data, hi.ctypes.data)
        return lo, hi

    def update(self, x, dfdx, gx, dgdx, xmin, xmax):
        x = np.array(x, dtype=np.float64)
        gx = np.ascontiguousarray(np.atleast_1d(gx), dtype=np.float64)
        dg = np.ascontiguousarray(np.concatenate([np.asarray(d, dtype=np.float64) for d in dgdx]))
        args = [np.ascontiguousarray(a, dtype=np.float64) for a in (dfdx, xmin, xmax)]
        self.l.mmaref_update(self.h, x.ctypes.data, args[0].ctypes.data, gx.ctypes.data, dg.ctypes.data,
                             args[1].ctypes.data, args[2].ctypes.data)
        return x

    def design_change(self, x, xold):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.l.mmaref_design_change(self.h, x.ctypes.data, xold.ctypes.data)

    def close(self):
        if self.h:
            self.l.mmaref_destroy(self.h)
            self.h = None
