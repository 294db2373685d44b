"""Host-side mirror of the reference's class surface over the C-ABI (tests / bench only).

Names, argument order and meaning follow the reference classes so parity tests read like the
reference's own driver (main.cc:31-102):

    TopOpt           -> Context (grid + partition; TopOpt.cc:146-308)
    LinearElasticity -> LinearElasticity.ComputeObjectiveConstraintsSensitivities(...)
    Filter           -> Filter.FilterProject / Gradients / IncreaseBeta / GetMND
    MMA              -> MMA.SetOuterMovelimit / Update / DesignChange

Vectors are numpy float64 arrays (host: staged inside the C call, the end-to-end path) or DevVec
(resident device memory owned by the library's allocator).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import GridDesc, SolverOpts, SolveStats, TobError, check


class DevVec:
    """A device-resident vector (doubles unless dtype given) allocated through tob_dev_alloc."""

    def __init__(self, ctx: "Context", n: int, dtype=np.float64):
        self.ctx, self.n, self.dtype = ctx, int(n), np.dtype(dtype)
        p = C.c_void_p()
        check(ctx.lib.tob_dev_alloc(ctx.h, self.n * self.dtype.itemsize, C.byref(p)))
        self.ptr = p.value

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.size == self.n
        check(self.ctx.lib.tob_memcpy(self.ctx.h, self.ptr, a.ctypes.data, a.nbytes))
        return self

    def download(self):
        out = np.empty(self.n, dtype=self.dtype)
        check(self.ctx.lib.tob_memcpy(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.tob_dev_free(self.ctx.h, self.ptr)
            self.ptr = None


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_lib.load().tob_nccl_unique_id(buf))
    return buf.raw


def _ptr(v):
    if v is None:
        return None
    if isinstance(v, DevVec):
        return v.ptr
    if isinstance(v, np.ndarray):
        assert v.flags["C_CONTIGUOUS"]
        return v.ctypes.data
    if isinstance(v, int):
        return v
    raise TypeError(type(v))


class Context:
    """Grid + z-slab partition + stream + NCCL communicator (replaces DM da_nodes / da_elem)."""

    def __init__(self, nxn, nyn, nzn, xc=(0.0, 2.0, 0.0, 1.0, 0.0, 1.0), rank=0, nranks=1, device=-1,
                 nccl_id: bytes | None = None, stream: int | None = None):
        self.lib = _lib.load()
        d = GridDesc()
        d.nxn, d.nyn, d.nzn = nxn, nyn, nzn
        for i in range(6):
            d.xc[i] = xc[i]
        d.rank, d.nranks, d.device = rank, nranks, device
        self._idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id else None
        d.nccl_id = C.cast(self._idbuf, C.c_void_p) if self._idbuf else None
        d.stream = stream
        h = C.c_void_p()
        check(self.lib.tob_ctx_create(C.byref(d), C.byref(h)))
        self.h = h
        a, b, c_, e = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_longlong()
        check(self.lib.tob_ctx_sizes(h, C.byref(a), C.byref(b), C.byref(c_), C.byref(e)))
        self.nel_owned, self.ndof_owned, self.nel, self.ndof = a.value, b.value, c_.value, e.value
        k = [C.c_int() for _ in range(4)]
        check(self.lib.tob_ctx_slab(h, *[C.byref(x) for x in k]))
        self.kn0, self.kn1, self.ke0, self.ke1 = [x.value for x in k]
        self.nxn, self.nyn, self.nzn = nxn, nyn, nzn
        self.rank, self.nranks = rank, nranks

    def sync(self):
        check(self.lib.tob_ctx_sync(self.h))

    def launch_count(self, reset=False):
        return int(self.lib.tob_ctx_launch_count(self.h, 1 if reset else 0))

    def stream(self):
        return self.lib.tob_ctx_stream(self.h)

    def dev(self, n, dtype=np.float64):
        return DevVec(self, n, dtype)

    def close(self):
        if self.h:
            self.lib.tob_ctx_destroy(self.h)
            self.h = None

    def scale(self, v, a, n=None):
        """VecScale (main.cc:73) on a host ndarray or DevVec."""
        n = int(n if n is not None else (v.n if isinstance(v, DevVec) else v.size))
        check(self.lib.tob_vec_scale(self.h, _ptr(v), a, n))

    def microbench_fp64(self, kind):
        v = C.c_double()
        check(self.lib.tob_microbench_fp64(self.h, kind, C.byref(v)))
        return v.value

    def microbench_copy(self, nbytes):
        v = C.c_double()
        check(self.lib.tob_microbench_copy(self.h, nbytes, C.byref(v)))
        return v.value


def solver_opts(**kw) -> SolverOpts:
    o = SolverOpts()
    _lib.load().tob_solver_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class LinearElasticity:
    """Mirror of class LinearElasticity (LinearElasticity.h:21-109)."""

    def __init__(self, ctx: Context, nu=0.3, opts: SolverOpts | None = None):
        self.ctx, self.lib = ctx, ctx.lib
        self.opts = opts if opts is not None else solver_opts()
        h = C.c_void_p()
        check(self.lib.tob_elast_create(ctx.h, nu, C.byref(self.opts), C.byref(h)))
        self.h = h

    def SetOpts(self, **kw):
        for k, v in kw.items():
            setattr(self.opts, k, v)
        check(self.lib.tob_elast_set_opts(self.h, C.byref(self.opts)))

    def SetBCLoad(self, fixed, rhs):
        fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        check(self.lib.tob_elast_set_bc_load(self.h, fixed.ctypes.data, rhs.ctypes.data))

    def GetKE(self):
        ke = np.empty(576)
        check(self.lib.tob_elast_get_ke(self.h, ke.ctypes.data))
        return ke.reshape(24, 24)

    def SetDensity(self, xPhys, Emin, Emax, penal):
        check(self.lib.tob_elast_set_density(self.h, _ptr(xPhys), Emin, Emax, penal))

    def SpMV(self, u, y=None):
        if y is None:
            y = np.empty(self.ctx.ndof_owned)
        check(self.lib.tob_elast_spmv(self.h, _ptr(u), _ptr(y)))
        return y

    def SpMVtf32(self, u, y=None):
        """single-precision tensor-core variant (tcgen05 kind::tf32); fp64 arrays at the boundary"""
        if y is None:
            y = np.empty(self.ctx.ndof_owned)
        check(self.lib.tob_elast_spmv_tf32(self.h, _ptr(u), _ptr(y)))
        return y

    def SpMVtf32Bench(self, reps=20):
        ms = C.c_double()
        check(self.lib.tob_elast_spmv_tf32_bench(self.h, reps, C.byref(ms)))
        return ms.value

    def Diag(self):
        d = np.empty(self.ctx.ndof_owned)
        check(self.lib.tob_elast_diag(self.h, d.ctypes.data))
        return d

    def SpMVBench(self, kernel=0, reps=20):
        ms = C.c_double()
        check(self.lib.tob_elast_spmv_bench(self.h, kernel, reps, C.byref(ms)))
        return ms.value

    def SolveState(self, xPhys, Emin, Emax, penal):
        check(self.lib.tob_elast_solve(self.h, _ptr(xPhys), Emin, Emax, penal))
        return self.stats()

    def ComputeObjectiveConstraintsSensitivities(self, dfdx, dgdx, xPhys, Emin, Emax, penal, volfrac):
        """Returns (fx, gx); dfdx/dgdx are written in place (LinearElasticity.cc:363-445)."""
        fx, gx = C.c_double(), C.c_double()
        check(self.lib.tob_elast_solve_obj_sens(self.h, _ptr(xPhys), Emin, Emax, penal, volfrac, C.byref(fx),
                                                C.byref(gx), _ptr(dfdx), _ptr(dgdx)))
        return fx.value, gx.value

    def ComputeObjectiveConstraints(self, xPhys, Emin, Emax, penal, volfrac):
        fx, gx = C.c_double(), C.c_double()
        check(self.lib.tob_elast_obj(self.h, _ptr(xPhys), Emin, Emax, penal, volfrac, C.byref(fx), C.byref(gx)))
        return fx.value, gx.value

    def ComputeSensitivities(self, dfdx, dgdx, xPhys, Emin, Emax, penal, volfrac):
        check(self.lib.tob_elast_sens(self.h, _ptr(xPhys), Emin, Emax, penal, volfrac, _ptr(dfdx), _ptr(dgdx)))

    def GetStateField(self):
        U = np.empty(self.ctx.ndof_owned)
        check(self.lib.tob_elast_get_state(self.h, U.ctypes.data))
        return U

    def SetStateField(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        check(self.lib.tob_elast_set_state(self.h, U.ctypes.data))

    def GetRHS(self):
        r = np.empty(self.ctx.ndof_owned)
        check(self.lib.tob_elast_get_rhs(self.h, r.ctypes.data))
        return r

    def GmgLevels(self):
        n = C.c_int()
        check(self.lib.tob_elast_gmg_nlevels(self.h, C.byref(n)))
        return n.value

    def GmgLevelMatrix(self, level):
        """None for a matrix-free level; for a Galerkin level the assembled operator as scipy CSR in the oracle's dof order
        (3 * node + c, node = i + nx * (j + ny * k))."""
        import scipy.sparse as sp
        dims, st = (C.c_int * 4)(), C.c_int()
        check(self.lib.tob_elast_gmg_level_stencil(self.h, level, None, dims, C.byref(st)))
        if not st.value:
            return None
        nx, ny, nz, px = list(dims)
        A = np.empty((243, nz, ny, px))
        check(self.lib.tob_elast_gmg_level_stencil(self.h, level, A.ctypes.data, dims, C.byref(st)))
        A = A[:, :, :, :nx].reshape(27, 3, 3, nz, ny, nx)
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        rows, cols, vals = [], [], []
        for o in range(27):
            di, dj, dk = o % 3 - 1, (o // 3) % 3 - 1, o // 9 - 1
            ok = (i + di >= 0) & (i + di < nx) & (j + dj >= 0) & (j + dj < ny) & (k + dk >= 0) & (k + dk < nz)
            n = (i + nx * (j + ny * k))[ok]
            m = ((i + di) + nx * ((j + dj) + ny * (k + dk)))[ok]
            for c in range(3):
                for c2 in range(3):
                    rows.append(3 * n + c); cols.append(3 * m + c2); vals.append(A[o, c, c2][ok])
        nd = 3 * nx * ny * nz
        return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nd, nd)).tocsr()

    def SetProfile(self, on=True):
        check(self.lib.tob_elast_set_profile(self.h, int(on)))

    def stats(self) -> SolveStats:
        s = SolveStats()
        check(self.lib.tob_elast_last_stats(self.h, C.byref(s)))
        return s

    def close(self):
        if self.h:
            self.lib.tob_elast_destroy(self.h)
            self.h = None


class Filter:
    """Mirror of class Filter (Filter.h:34-92) including the PDE filter it owns."""

    def __init__(self, ctx: Context, filterT: int, Rin: float):
        self.ctx, self.lib = ctx, ctx.lib
        h = C.c_void_p()
        check(self.lib.tob_filter_create(ctx.h, filterT, Rin, C.byref(h)))
        self.h = h

    def ElemConn(self):
        v = C.c_int()
        check(self.lib.tob_filter_elemconn(self.h, C.byref(v)))
        return v.value

    def FilterProject(self, x, xTilde, xPhys, projectionFilter, beta, eta):
        check(self.lib.tob_filter_project(self.h, _ptr(x), _ptr(xTilde), _ptr(xPhys), int(bool(projectionFilter)),
                                          beta, eta))

    def Gradients(self, x, xTilde, dfdx, m, dgdx, projectionFilter, beta, eta):
        arr = (C.c_void_p * m)(*[_ptr(d) for d in dgdx])
        check(self.lib.tob_filter_gradients(self.h, _ptr(x), _ptr(xTilde), _ptr(dfdx), m, arr,
                                            int(bool(projectionFilter)), beta, eta))

    def IncreaseBeta(self, beta, betaFinal, gx, itr, ch):
        b, ch_flag = C.c_double(beta), C.c_int()
        check(self.lib.tob_filter_increase_beta(C.byref(b), betaFinal, gx, itr, ch, C.byref(ch_flag)))
        return b.value, bool(ch_flag.value)

    def GetMND(self, xPhys):
        v = C.c_double()
        check(self.lib.tob_filter_mnd(self.h, _ptr(xPhys), C.byref(v)))
        return v.value

    def GetHs(self):
        hs = np.empty(self.ctx.nel_owned)
        check(self.lib.tob_filter_get_hs(self.h, hs.ctypes.data))
        return hs

    def PDEApply(self, x, out=None):
        if out is None:
            out = np.empty(self.ctx.nel_owned)
        check(self.lib.tob_filter_pde_apply(self.h, _ptr(x), _ptr(out)))
        return out

    def stats(self) -> SolveStats:
        s = SolveStats()
        check(self.lib.tob_filter_last_stats(self.h, C.byref(s)))
        return s

    def close(self):
        if self.h:
            self.lib.tob_filter_destroy(self.h)
            self.h = None


class MMA:
    """Mirror of class MMA (MMA.h:29-140), device implementation."""

    def __init__(self, ctx: Context, m: int, a=None, c=None, d=None):
        self.ctx, self.lib, self.m = ctx, ctx.lib, m

        def arr(v):
            return None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        a, c, d = arr(a), arr(c), arr(d)
        h = C.c_void_p()
        check(self.lib.tob_mma_create(ctx.h, m, _ptr(a), _ptr(c), _ptr(d), C.byref(h)))
        self.h = h

    def SetOuterMovelimit(self, Xmin, Xmax, movlim, x, xmin, xmax):
        check(self.lib.tob_mma_set_outer_movelimit(self.h, Xmin, Xmax, movlim, _ptr(x), _ptr(xmin), _ptr(xmax)))

    def Update(self, xval, dfdx, gx, dgdx, xmin, xmax):
        gxa = np.ascontiguousarray(np.atleast_1d(gx), dtype=np.float64)
        arr = (C.c_void_p * self.m)(*[_ptr(d) for d in dgdx])
        check(self.lib.tob_mma_update(self.h, _ptr(xval), _ptr(dfdx), gxa.ctypes.data, arr, _ptr(xmin), _ptr(xmax)))

    def DesignChange(self, x, xold):
        v = C.c_double()
        check(self.lib.tob_mma_design_change(self.h, _ptr(x), _ptr(xold), C.byref(v)))
        return v.value

    def Restart(self):
        n = self.ctx.nel_owned
        out = [np.empty(n) for _ in range(4)]
        k = C.c_int()
        check(self.lib.tob_mma_get_restart(self.h, *[o.ctypes.data for o in out], C.byref(k)))
        return (*out, k.value)

    def SetRestart(self, xo1, xo2, U, L, k):
        arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in (xo1, xo2, U, L)]
        check(self.lib.tob_mma_set_restart(self.h, *[a.ctypes.data for a in arrs], k))

    def close(self):
        if self.h:
            self.lib.tob_mma_destroy(self.h)
            self.h = None
