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
        l.cref_mg_setup.restype = C.c_double
        l.cref_mg_setup.argtypes = [C.c_void_p, C.c_int]
        l.cref_fgmres_mg.restype = C.c_int
        l.cref_fgmres_mg.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        l.cref_mg_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        for f in ("cref_level_nnz", "cref_level_ndof"):
            getattr(l, f).restype = C.c_longlong
            getattr(l, f).argtypes = [C.c_void_p, C.c_int]
        l.cref_level_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        l.cref_mg_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.cref_cone_filter.restype = C.c_double
        l.cref_cone_filter.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
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

    # ---- the reference's solver configuration: FGMRES + PCMG(Galerkin) + GMRES/SOR smoothers -------------
    def mg_setup(self, nlvls=4):
        """KSPSetUp after an operator change (LinearElasticity.cc:198-200): Galerkin products of every level."""
        sec = self.l.cref_mg_setup(self.h, nlvls)
        if sec < 0:
            raise ValueError("mesh dimension not compatible with the number of multigrid levels")
        return sec

    def fgmres_mg(self, U=None, rtol=1e-5, maxits=200, restart=100):
        U = np.zeros(self.ndof) if U is None else np.ascontiguousarray(U, dtype=np.float64)
        rr, sec = C.c_double(), C.c_double()
        its = self.l.cref_fgmres_mg(self.h, U.ctypes.data, rtol, maxits, restart, C.byref(rr), C.byref(sec))
        return U, its, rr.value, sec.value

    def mg_apply(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.ndof)
        self.l.cref_mg_apply(self.h, b.ctypes.data, x.ctypes.data)
        return x

    def level_matrix(self, l):
        import scipy.sparse as sp
        n, nnz = self.l.cref_level_ndof(self.h, l), self.l.cref_level_nnz(self.h, l)
        rp, col, val = np.empty(n + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz)
        self.l.cref_level_csr(self.h, l, rp.ctypes.data, col.ctypes.data, val.ctypes.data)
        return sp.csr_matrix((val, col, rp), shape=(n, n))

    def cone_filter(self, rmin, x, mode=0):
        """mode 0: (H x)/Hs (Filter.cc:66-71); mode 1: H (x/Hs) (Filter.cc:178-192)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(self.nel)
        sec = self.l.cref_cone_filter(self.h, rmin, mode, x.ctypes.data, out.ctypes.data)
        return out, sec

    def close(self):
        if self.h:
            self.l.cref_destroy(self.h)
            self.h = None
