"""ctypes wrapper of oracle/c/libcpu_ref.so (C/OpenMP restatement of the reference's CPU state-solve
path).  TEST / BASELINE INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(os.path.join(_HERE, "c", "libcpu_ref.so"))
        l.cref_create.restype = C.c_void_p
        l.cref_create.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_int]
        l.cref_destroy.argtypes = [C.c_void_p]
        for f in ("cref_ndof", "cref_nel", "cref_nnz"):
            getattr(l, f).restype = C.c_longlong
            getattr(l, f).argtypes = [C.c_void_p]
        l.cref_get_ke.argtypes = [C.c_void_p, C.c_void_p]
        l.cref_get_rhs.argtypes = [C.c_void_p, C.c_void_p]
        l.cref_assemble.restype = C.c_double
        l.cref_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]
        l.cref_spmv.restype = C.c_double
        l.cref_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.cref_pcg.restype = C.c_int
        l.cref_pcg.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        l.cref_objective.restype = C.c_double
        l.cref_objective.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_double] * 4 + [C.POINTER(C.c_double)] * 2 + [C.c_void_p]
        _lib = l
    return _lib


class CpuRef:
    def __init__(self, nxn, nyn, nzn, xc=(0, 2, 0, 1, 0, 1), nu=0.3, nthreads=0):
        self.l = lib()
        dx, dy, dz = (xc[1] - xc[0]) / (nxn - 1), (xc[3] - xc[2]) / (nyn - 1), (xc[5] - xc[4]) / (nzn - 1)
        self.h = self.l.cref_create(nxn, nyn, nzn, dx, dy, dz, nu, nthreads)
        self.ndof, self.nel, self.nnz = self.l.cref_ndof(self.h), self.l.cref_nel(self.h), self.l.cref_nnz(self.h)

    def ke(self):
        k = np.empty(576)
        self.l.cref_get_ke(self.h, k.ctypes.data)
        return k.reshape(24, 24)

    def rhs(self):
        r = np.empty(self.ndof)
        self.l.cref_get_rhs(self.h, r.ctypes.data)
        return r

    def assemble(self, x, Emin=1e-9, Emax=1.0, penal=3.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.l.cref_assemble(self.h, x.ctypes.data, Emin, Emax, penal)

    def spmv(self, u, reps=1):
        u = np.ascontiguousarray(u, dtype=np.float64)
        y = np.empty(self.ndof)
        sec = self.l.cref_spmv(self.h, u.ctypes.data, y.ctypes.data, reps)
        return y, sec

    def pcg(self, U=None, rtol=1e-5, maxit=100000):
        U = np.zeros(self.ndof) if U is None else np.ascontiguousarray(U, dtype=np.float64)
        rr, sec = C.c_double(), C.c_double()
        its = self.l.cref_pcg(self.h, U.ctypes.data, rtol, maxit, C.byref(rr), C.byref(sec))
        return U, its, rr.value, sec.value

    def objective(self, U, x, Emin=1e-9, Emax=1.0, penal=3.0, volfrac=0.12):
        x = np.ascontiguousarray(x, dtype=np.float64)
        fx, gx = C.c_double(), C.c_double()
        df = np.empty(self.nel)
        sec = self.l.cref_objective(self.h, U.ctypes.data, x.ctypes.data, Emin, Emax, penal, volfrac, C.byref(fx),
                                    C.byref(gx), df.ctypes.data)
        return fx.value, gx.value, df, sec

    def close(self):
        if self.h:
            self.l.cref_destroy(self.h)
            self.h = None
