"""CPU oracle for the TopOpt_in_PETSc hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy/scipy restatement of the reference's arithmetic for the
state solve, objective/sensitivities and the design filters.  It is imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg -- the
product path (topopt_in_petsc_b200/ + libtopopt_b200.so) never touches it.

PARITY STATUS: pinned by the reference's own code.  The reference ships no tests or
fixtures and its build needs PETSc + MPI (absent here), but its seven source files
compile unmodified, where they lie, against oracle/petsc_lite -- a serial restatement
of the PETSc/MPI slice they call (KSPSolve replaced by an exact Jacobi-PCG solve) --
into oracle/_ref/topopt_ref (oracle/Makefile).  tests/test_reference_cpu.py checks this
oracle's design loop against that program's output (tests/golden/reference_runs.json:
printed history to the printed 6 decimals, final design field to 1e-10) for the
density, sensitivity and PDE filters and the Heaviside projection.  Also pinned by
(i) the surveyor's known answers (tests/golden/appendix_c.json),
(ii) the independent C restatement (oracle/c/), (iii) the reference's unmodified MMA.cc
as a library (oracle/_ref/libmma_ref.so, tests/test_mma_ref.py).

Every function cites the reference file:line (paths relative to the reference
checkout) it restates.  Numbering conventions:
  node id  = i + nxn*(j + nyn*k)   (i fastest),  dof = 3*node + c
  elem id  = i + nex*(j + ney*k)
  element-local node order (i,j,k),(i+1,j,k),(i+1,j+1,k),(i,j+1,k), then k+1
  (LinearElasticity.cc:819-826).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------
# element matrices
# --------------------------------------------------------------------------

_GP = 0.577350269189626  # LinearElasticity.cc:898 (truncated sqrt(1/3), kept verbatim)


def _dshape(xi, eta, zeta):
    """Differentiated trilinear shape functions, LinearElasticity.cc:1009-1039."""
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1], dtype=float)
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1], dtype=float)
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1], dtype=float)
    dxi = 0.125 * sx * (1 + sy * eta) * (1 + sz * zeta)
    deta = 0.125 * sy * (1 + sx * xi) * (1 + sz * zeta)
    dzeta = 0.125 * sz * (1 + sx * xi) * (1 + sy * eta)
    return dxi, deta, dzeta


def hex8_ke(dx: float, dy: float, dz: float, nu: float) -> np.ndarray:
    """24x24 element stiffness for E=1, full 2x2x2 Gauss integration.

    Restates LinearElasticity::Hex8Isoparametric, LinearElasticity.cc:841-998,
    called with X,Y,Z of LinearElasticity.cc:118-120.  Row-major ke[j+24*i].
    """
    X = np.array([0, dx, dx, 0, 0, dx, dx, 0], dtype=float)
    Y = np.array([0, 0, dy, dy, 0, 0, dy, dy], dtype=float)
    Z = np.array([0, 0, 0, 0, dz, dz, dz, dz], dtype=float)
    lam = nu / ((1.0 + nu) * (1.0 - 2.0 * nu))  # :887
    mu = 1.0 / (2.0 * (1.0 + nu))  # :888
    C = np.zeros((6, 6))
    C[:3, :3] = lam
    C[0, 0] = C[1, 1] = C[2, 2] = lam + 2.0 * mu
    C[3, 3] = C[4, 4] = C[5, 5] = mu
    ke = np.zeros((24, 24))
    for xi in (-_GP, _GP):
        for eta in (-_GP, _GP):
            for zeta in (-_GP, _GP):
                dxi, deta, dzeta = _dshape(xi, eta, zeta)
                J = np.array([[dxi @ X, dxi @ Y, dxi @ Z],
                              [deta @ X, deta @ Y, deta @ Z],
                              [dzeta @ X, dzeta @ Y, dzeta @ Z]])
                detJ = np.linalg.det(J)
                invJ = np.linalg.inv(J)
                # physical derivatives dN/dx_a = sum_l invJ[a][l] dN/dxi_l  (:958-982)
                dN = invJ @ np.vstack([dxi, deta, dzeta])  # 3 x 8
                B = np.zeros((6, 24))
                for n in range(8):
                    gx, gy, gz = dN[:, n]
                    B[0, 3 * n + 0] = gx
                    B[1, 3 * n + 1] = gy
                    B[2, 3 * n + 2] = gz
                    B[3, 3 * n + 0] = gy
                    B[3, 3 * n + 1] = gx
                    B[4, 3 * n + 1] = gz
                    B[4, 3 * n + 2] = gy
                    B[5, 3 * n + 0] = gz
                    B[5, 3 * n + 2] = gx
                ke += detJ * (B.T @ C @ B)  # weights are 1 (:900)
    return ke


def pde_filter_ke(dx: float, dy: float, dz: float, R: float) -> np.ndarray:
    """8x8 Helmholtz element matrix KF = R^2 int grad N.grad N + int N N.

    Closed form of PDEFilt::PDEFilterMatrix, PDEFilter.cc:472-565 (the t3..t53
    expressions evaluate exactly these integrals; verified against 2x2x2
    quadrature in tests).  Local node order = the hex order above.
    """
    t3 = 1.0 / dx / dy
    t4 = 1.0 / dz
    t5 = R * R
    t6 = dx * dx
    t7 = t5 * t6
    t8 = dy * dy
    t9 = t7 * t8
    t10 = 3.0 * t9
    t11 = dz * dz
    t12 = t7 * t11
    t13 = 3.0 * t12
    t15 = t5 * t8 * t11
    t16 = 3.0 * t15
    t18 = t6 * t8 * t11
    t22 = t3 * t4 * (t10 + t13 + t16 + t18) / 27.0
    t23 = 6.0 * t15
    t27 = t3 * t4 * (t10 + t13 - t23 + t18) / 54.0
    t28 = 6.0 * t12
    t32 = t3 * t4 * (t10 - t28 - t23 + t18) / 108.0
    t36 = t3 * t4 * (t10 - t28 + t16 + t18) / 54.0
    t37 = 6.0 * t9
    t41 = t3 * t4 * (t37 - t13 - t16 - t18) / 54.0
    t45 = t3 * t4 * (t37 - t13 + t23 - t18) / 108.0
    t49 = t3 * t4 * (t37 + t28 + t23 - t18) / 216.0
    t53 = t3 * t4 * (t37 + t28 - t16 - t18) / 108.0
    a, b, c, d = t22, t27, t32, t36
    e, f, g, h = -t41, -t45, -t49, -t53
    # Row pattern of PDEFilter.cc:502-565.
    K = np.array([
        [a, b, c, d, e, f, g, h],
        [b, a, d, c, f, e, h, g],
        [c, d, a, b, g, h, e, f],
        [d, c, b, a, h, g, f, e],
        [e, f, g, h, a, b, c, d],
        [f, e, h, g, b, a, d, c],
        [g, h, e, f, c, d, a, b],
        [h, g, f, e, d, c, b, a],
    ])
    return K


# --------------------------------------------------------------------------
# grid / connectivity
# --------------------------------------------------------------------------

@dataclass
class Grid:
    """Structured hex grid. nxn,nyn,nzn are NODE counts (the reference's -nx -ny -nz)."""
    nxn: int
    nyn: int
    nzn: int
    xc: tuple = (0.0, 2.0, 0.0, 1.0, 0.0, 1.0)  # TopOpt.cc:109-114
    nu: float = 0.3

    def __post_init__(self):
        self.nex, self.ney, self.nez = self.nxn - 1, self.nyn - 1, self.nzn - 1
        self.nnode = self.nxn * self.nyn * self.nzn
        self.ndof = 3 * self.nnode  # TopOpt.cc:172
        self.nel = self.nex * self.ney * self.nez  # TopOpt.cc:173-175
        self.dx = (self.xc[1] - self.xc[0]) / self.nex  # TopOpt.cc:225-227
        self.dy = (self.xc[3] - self.xc[2]) / self.ney
        self.dz = (self.xc[5] - self.xc[4]) / self.nez

    def node_id(self, i, j, k):
        return i + self.nxn * (j + self.nyn * k)

    def elem_nodes(self) -> np.ndarray:
        """(nel, 8) global node ids, order of LinearElasticity.cc:819-826."""
        i, j, k = np.meshgrid(np.arange(self.nex), np.arange(self.ney), np.arange(self.nez), indexing="ij")
        i = i.transpose(2, 1, 0).ravel()
        j = j.transpose(2, 1, 0).ravel()
        k = k.transpose(2, 1, 0).ravel()
        n = self.node_id
        return np.stack([n(i, j, k), n(i + 1, j, k), n(i + 1, j + 1, k), n(i, j + 1, k),
                         n(i, j, k + 1), n(i + 1, j, k + 1), n(i + 1, j + 1, k + 1), n(i, j + 1, k + 1)], axis=1)

    def edof(self) -> np.ndarray:
        en = self.elem_nodes()
        return (3 * en[:, :, None] + np.arange(3)[None, None, :]).reshape(self.nel, 24)


def cantilever_bc_load(g: Grid):
    """N (1 free / 0 clamped) and RHS, LinearElasticity.cc:143-171.

    Wall x = xmin fully clamped; line load in z of -0.001 per node on the edge
    (x = xmax, z = zmin), halved at the two y-end nodes.
    """
    N = np.ones(g.ndof)
    RHS = np.zeros(g.ndof)
    load = -0.001
    i0 = 0
    for k in range(g.nzn):
        for j in range(g.nyn):
            n = g.node_id(i0, j, k)
            N[3 * n:3 * n + 3] = 0.0
    for j in range(g.nyn):
        n = g.node_id(g.nxn - 1, j, 0)
        RHS[3 * n + 2] = load / 2.0 if j in (0, g.nyn - 1) else load
    return N, RHS


def simp_modulus(xphys, Emin, Emax, penal):
    """E_e = Emin + x^p (Emax-Emin), LinearElasticity.cc:519."""
    return Emin + np.power(xphys, penal) * (Emax - Emin)


def assemble_K(g: Grid, KE, xphys, Emin, Emax, penal, N):
    """Assembled global stiffness with Dirichlet rows, LinearElasticity.cc:487-549.

    K = sum_e E_e KE scattered by edof (:510-526); K <- N K N + (I - N) (:532-538).
    """
    ed = g.edof()
    E = simp_modulus(xphys, Emin, Emax, penal)
    rows = np.repeat(ed, 24, axis=1).ravel()
    cols = np.tile(ed, (1, 24)).ravel()
    vals = (E[:, None] * KE.ravel()[None, :]).ravel()
    K = sp.coo_matrix((vals, (rows, cols)), shape=(g.ndof, g.ndof)).tocsr()
    D = sp.diags(N)
    K = D @ K @ D + sp.diags(1.0 - N)
    return K.tocsr()


def spmv_matrix_free(g: Grid, KE, E, N, u):
    """Reference-semantics operator A u = N.(K (N.u)) + (1-N).u without assembling."""
    ed = g.edof()
    um = N * u
    ue = um[ed]  # nel x 24
    ye = (ue @ KE.T) * E[:, None]
    y = np.zeros(g.ndof)
    np.add.at(y, ed.ravel(), ye.ravel())
    return N * y + (1.0 - N) * u


def solve_state(K, rhs, tol=1e-12, x0=None, maxiter=20000):
    """Tight Jacobi-PCG solve of K U = N.RHS (reference: FGMRES+PCMG to 1e-5, LinearElasticity.cc:621-651)."""
    d = K.diagonal()
    M = spla.LinearOperator(K.shape, matvec=lambda r: r / d)
    its = [0]

    def cb(_):
        its[0] += 1
    U, info = spla.cg(K, rhs, x0=x0, rtol=tol, atol=0.0, M=M, maxiter=maxiter, callback=cb)
    return U, its[0], info


def objective_sens(g: Grid, KE, U, xphys, Emin, Emax, penal, volfrac):
    """fx, gx, dfdx, dgdx, LinearElasticity.cc:403-437."""
    ed = g.edof()
    ue = U[ed]
    uKu = np.einsum("ei,ij,ej->e", ue, KE, ue)
    E = simp_modulus(xphys, Emin, Emax, penal)
    fx = float(np.sum(E * uKu))
    dfdx = -penal * np.power(xphys, penal - 1.0) * (Emax - Emin) * uKu
    gx = float(np.sum(xphys) / g.nel - volfrac)
    dgdx = np.full(g.nel, 1.0 / g.nel)
    return fx, gx, dfdx, dgdx


# --------------------------------------------------------------------------
# cone filter (Filter.cc)
# --------------------------------------------------------------------------

def filter_elemconn(g: Grid, R: float) -> int:
    """ElemConn, Filter.cc:326-327 (integer division on NODE counts)."""
    ec = int(max(math.ceil(R / g.dx) - 1, max(math.ceil(R / g.dy) - 1, math.ceil(R / g.dz) - 1)))
    return min(ec, min((g.nxn - 1) // 2, min((g.nyn - 1) // 2, (g.nzn - 1) // 2)))


def filter_matrix(g: Grid, R: float):
    """H (csr, nel x nel) and Hs = H 1, Filter.cc:375-448.

    Element-centre coordinates are the uniform DMDA coordinates between d/2 and
    L - d/2 (:375-379); weight R - dist for dist < R inside the clipped box
    |di|,|dj|,|dk| <= ElemConn (:404-440).
    """
    s = filter_elemconn(g, R)
    nex, ney, nez = g.nex, g.ney, g.nez

    def coords(n, d):
        lo, hi = d / 2.0, n * d - d / 2.0
        if n == 1:
            return np.array([lo])
        return lo + np.arange(n) * ((hi - lo) / (n - 1))
    cx, cy, cz = coords(nex, g.dx), coords(ney, g.dy), coords(nez, g.dz)
    I, J, Kk = np.meshgrid(np.arange(nex), np.arange(ney), np.arange(nez), indexing="ij")
    I = I.transpose(2, 1, 0).ravel(); J = J.transpose(2, 1, 0).ravel(); Kk = Kk.transpose(2, 1, 0).ravel()
    rows, cols, vals = [], [], []
    eid = I + nex * (J + ney * Kk)
    for dk in range(-s, s + 1):
        for dj in range(-s, s + 1):
            for di in range(-s, s + 1):
                i2, j2, k2 = I + di, J + dj, Kk + dk
                ok = (i2 >= 0) & (i2 < nex) & (j2 >= 0) & (j2 < ney) & (k2 >= 0) & (k2 < nez)
                i2c, j2c, k2c = i2[ok], j2[ok], k2[ok]
                dist = np.sqrt((cx[I[ok]] - cx[i2c]) ** 2 + (cy[J[ok]] - cy[j2c]) ** 2 + (cz[Kk[ok]] - cz[k2c]) ** 2)
                keep = dist < R
                rows.append(eid[ok][keep])
                cols.append((i2c + nex * (j2c + ney * k2c))[keep])
                vals.append((R - dist)[keep])
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(g.nel, g.nel)).tocsr()
    Hs = np.asarray(H.sum(axis=1)).ravel()
    return H, Hs, s


def smooth_projection(x, beta, eta):
    """Filter.h:80-83."""
    return (np.tanh(beta * eta) + np.tanh(beta * (x - eta))) / (np.tanh(beta * eta) + np.tanh(beta * (1.0 - eta)))


def chainrule_smooth_projection(x, beta, eta):
    """Filter.h:85-88."""
    return beta * (1.0 - np.tanh(beta * (x - eta)) ** 2) / (np.tanh(beta * eta) + np.tanh(beta * (1.0 - eta)))


def mnd(x):
    """Filter.cc:206-225."""
    return float(np.sum(4.0 * x * (1.0 - x)) / x.size)


def increase_beta(beta, beta_final, gx, itr, ch):
    """Filter.cc:268-288. Returns (new beta, changeBeta flag)."""
    change = False
    if (ch < 0.01 or itr % 10 == 0) and beta < beta_final and gx < 0.000001:
        change = True
        beta = beta + 1 if beta < 7 else beta * 1.2
        if beta > beta_final:
            beta = beta_final
            change = False
    return beta, change


@dataclass
class PDEFilterOracle:
    """PDEFilt, PDEFilter.cc:28-267: K = sum KF (Neumann), T[node,elem] = 1/8."""
    g: Grid
    rmin: float
    K: sp.csr_matrix = field(init=False)
    T: sp.csr_matrix = field(init=False)

    def __post_init__(self):
        g = self.g
        self.R = self.rmin / 2.0 / math.sqrt(3)  # PDEFilter.cc:30
        KF = pde_filter_ke(g.dx, g.dy, g.dz, self.R)
        self.KF = KF
        en = g.elem_nodes()
        rows = np.repeat(en, 8, axis=1).ravel()
        cols = np.tile(en, (1, 8)).ravel()
        vals = np.tile(KF.ravel(), g.nel)
        self.K = sp.coo_matrix((vals, (rows, cols)), shape=(g.nnode, g.nnode)).tocsr()
        self.T = sp.coo_matrix((np.full(8 * g.nel, 0.125), (en.ravel(), np.repeat(np.arange(g.nel), 8))),
                               shape=(g.nnode, g.nel)).tocsr()
        self.elem_vol = g.dx * g.dy * g.dz  # PDEFilter.cc:65
        self._lu = None

    def apply(self, x, tol=1e-13):
        """FilterProject (= Gradients), PDEFilter.cc:189-218: T^T K^-1 (vol T x)."""
        rhs = self.T @ x
        u0 = rhs.copy()
        rhs = rhs * self.elem_vol
        d = self.K.diagonal()
        M = spla.LinearOperator(self.K.shape, matvec=lambda r: r / d)
        U, info = spla.cg(self.K, rhs, x0=u0, rtol=tol, atol=0.0, M=M, maxiter=5000)
        assert info == 0
        return self.T.T @ U


@dataclass
class FilterOracle:
    """Filter::FilterProject / Gradients, Filter.cc:60-204."""
    g: Grid
    filter_type: int
    rmin: float

    def __post_init__(self):
        self.H = self.Hs = self.pde = None
        if self.filter_type in (0, 1):
            self.H, self.Hs, self.elemconn = filter_matrix(self.g, self.rmin)
        elif self.filter_type == 2:
            self.pde = PDEFilterOracle(self.g, self.rmin)

    def filter_project(self, x, projection, beta, eta):
        if self.filter_type == 1:
            xt = (self.H @ x) / self.Hs  # :66-71
        elif self.filter_type == 2:
            xt = np.clip(self.pde.apply(x), 0.0, 1.0)  # :73-102
        else:
            xt = x.copy()  # :104-107
        xp = smooth_projection(xt, beta, eta) if projection else xt.copy()  # :110-114
        return xt, xp

    def gradients(self, x, xtilde, dfdx, dgdx_list, projection, beta, eta):
        dfdx = dfdx.copy()
        dgdx_list = [d.copy() for d in dgdx_list]
        if projection:  # :125-164
            dxp = chainrule_smooth_projection(xtilde, beta, eta)
            dfdx *= dxp
            dgdx_list = [d * dxp for d in dgdx_list]
        if self.filter_type == 0:  # :167-177 (only dfdx)
            dfdx = (self.H @ (dfdx * x)) / self.Hs / x
        elif self.filter_type == 1:  # :178-192
            dfdx = self.H @ (dfdx / self.Hs)
            dgdx_list = [self.H @ (d / self.Hs) for d in dgdx_list]
        elif self.filter_type == 2:  # :193-200
            dfdx = self.pde.apply(dfdx)
            dgdx_list = [self.pde.apply(d) for d in dgdx_list]
        return dfdx, dgdx_list


# --------------------------------------------------------------------------
# multigrid transfer (DMDA Q1 interpolation, LinearElasticity.cc:689-707)
# --------------------------------------------------------------------------

def q1_prolongation_1d(nc: int) -> sp.csr_matrix:
    """1-D linear interpolation from nc coarse nodes to 2*nc-1 fine nodes."""
    nf = 2 * nc - 1
    rows, cols, vals = [], [], []
    for f in range(nf):
        if f % 2 == 0:
            rows.append(f); cols.append(f // 2); vals.append(1.0)
        else:
            rows += [f, f]; cols += [f // 2, f // 2 + 1]; vals += [0.5, 0.5]
    return sp.coo_matrix((vals, (rows, cols)), shape=(nf, nc)).tocsr()


def q1_prolongation(ncx, ncy, ncz, ndof=3) -> sp.csr_matrix:
    """Trilinear prolongation per dof component, node-major/dof-minor ordering."""
    Px, Py, Pz = q1_prolongation_1d(ncx), q1_prolongation_1d(ncy), q1_prolongation_1d(ncz)
    P = sp.kron(Pz, sp.kron(Py, Px))
    if ndof > 1:
        P = sp.kron(P, sp.identity(ndof))
    return P.tocsr()


# --------------------------------------------------------------------------
# MMA (restatement of MMA.cc, single rank) -- cross-checked against the
# reference's own MMA.cc built into oracle/_ref (see tests/test_mma_ref.py)
# --------------------------------------------------------------------------

class MMAOracle:
    """MMA.cc:293-330 (ctor defaults), :386-426, :499-946."""

    def __init__(self, n, m, a=None, c=None, d=None):
        self.n, self.m, self.k = n, m, 0
        self.asyminit, self.asymdec, self.asyminc = 0.5, 0.7, 1.2
        self.a = np.zeros(m) if a is None else np.array(a, float)
        self.c = np.full(m, 1000.0) if c is None else np.array(c, float)
        self.d = np.zeros(m) if d is None else np.array(d, float)
        self.L = np.zeros(n); self.U = np.zeros(n)
        self.xo1 = np.zeros(n); self.xo2 = np.zeros(n)
        self.lam = np.zeros(m); self.mu = np.zeros(m); self.y = np.zeros(m); self.z = 0.0

    @staticmethod
    def set_outer_movelimit(Xmin, Xmax, movlim, x):
        """MMA.cc:386-405."""
        return np.maximum(Xmin, x - movlim), np.minimum(Xmax, x + movlim)

    def update(self, x, dfdx, gx, dgdx, xmin, xmax):
        """MMA::Update, MMA.cc:499-518. Returns the new x (x is not modified)."""
        self._gensub(x, dfdx, np.atleast_1d(gx), dgdx, xmin, xmax)
        self.xo2 = self.xo1.copy()
        self.xo1 = x.copy()
        return self._solve_dip()

    def _gensub(self, x, dfdx, gx, dgdx, xmin, xmax):
        """MMA.cc:522-649."""
        self.k += 1
        L, U = self.L, self.U
        if self.k < 3:
            L[:] = x - self.asyminit * (xmax - xmin)
            U[:] = x + self.asyminit * (xmax - xmin)
        else:
            h = (x - self.xo1) * (self.xo1 - self.xo2)
            gamma = np.where(h < 0, self.asymdec, np.where(h > 0, self.asyminc, 1.0))
            L[:] = x - gamma * (self.xo1 - L)
            U[:] = x + gamma * (U - self.xo1)
            xmi = np.maximum(1.0e-5, xmax - xmin)
            L[:] = np.minimum(np.maximum(L, x - 10.0 * xmi), x - 0.01 * xmi)
            U[:] = np.minimum(np.maximum(U, x + 0.01 * xmi), x + 10.0 * xmi)
        self.alpha = np.maximum(xmin, 0.9 * L + 0.1 * x)
        self.beta = np.minimum(xmax, 0.9 * U + 0.1 * x)
        feps = 1.0e-6
        dp, dm = np.maximum(0.0, dfdx), np.maximum(0.0, -dfdx)
        self.p0 = (U - x) ** 2 * (dp + 0.001 * np.abs(dfdx) + 0.5 * feps / (U - L))
        self.q0 = (x - L) ** 2 * (dm + 0.001 * np.abs(dfdx) + 0.5 * feps / (U - L))
        self.pij = [(U - x) ** 2 * np.maximum(0.0, dg) for dg in dgdx]
        self.qij = [(x - L) ** 2 * np.maximum(0.0, -dg) for dg in dgdx]
        self.b = np.array([np.sum(p / (U - x) + q / (x - L)) for p, q in zip(self.pij, self.qij)]) - gx

    def _xyz(self):
        """MMA.cc:690-740."""
        self.lam = np.maximum(self.lam, 0.0)
        self.y = np.maximum(0.0, self.lam - self.c)
        self.z = max(0.0, 10.0 * (float(self.lam @ self.a) - 1.0))
        pj, qj = self.p0.copy(), self.q0.copy()
        for j in range(self.m):
            pj += self.pij[j] * self.lam[j]
            qj += self.qij[j] * self.lam[j]
        x = (np.sqrt(pj) * self.L + np.sqrt(qj) * self.U) / (np.sqrt(pj) + np.sqrt(qj))
        return np.minimum(np.maximum(x, self.alpha), self.beta), pj, qj

    def _sums(self, x):
        return np.array([np.sum(p / (self.U - x) + q / (x - self.L)) for p, q in zip(self.pij, self.qij)])

    def _solve_dip(self):
        """MMA.cc:651-688 with DualGrad :742-777, DualHess :779-880, line search :882-900, residual :902-946."""
        m = self.m
        self.lam = self.c / 2.0
        self.mu = np.ones(m)
        tol = 1.0e-9 * math.sqrt(m + self.n)
        epsi, err = 1.0, 1.0
        x = None
        while epsi > tol:
            loop = 0
            while err > 0.9 * epsi and loop < 100:
                loop += 1
                x, pj, qj = self._xyz()
                grad = self._sums(x) - self.b - self.a * self.z - self.y
                grad = -grad - epsi / self.lam
                # Hessian
                PQ = np.stack([p / (self.U - x) ** 2 - q / (x - self.L) ** 2 for p, q in zip(self.pij, self.qij)], 1)
                df2 = -1.0 / (2.0 * pj / (self.U - x) ** 3 + 2.0 * qj / (x - self.L) ** 3)
                xp = (np.sqrt(pj) * self.L + np.sqrt(qj) * self.U) / (np.sqrt(pj) + np.sqrt(qj))
                df2[(xp < self.alpha) | (xp > self.beta)] = 0.0
                Hess = PQ.T @ (PQ * df2[:, None])
                self.lam = np.maximum(self.lam, 0.0)
                lamai = float(self.lam @ self.a)
                for j in range(m):
                    if self.lam[j] > self.c[j]:
                        Hess[j, j] += -1.0
                    Hess[j, j] += -self.mu[j] / self.lam[j]
                if lamai > 0.0:
                    Hess += -10.0 * np.outer(self.a, self.a)
                corr = 1e-4 * np.trace(Hess) / m
                if -corr < 1.0e-7:
                    corr = -1.0e-7
                Hess += corr * np.eye(m)
                s_lam = np.linalg.solve(Hess, grad)
                s_mu = -self.mu + epsi / self.lam - s_lam * self.mu / self.lam
                theta = 1.005
                theta = max(theta, np.max(-1.01 * s_lam / self.lam), np.max(-1.01 * s_mu / self.mu))
                theta = 1.0 / theta
                self.lam = self.lam + theta * s_lam
                self.mu = self.mu + theta * s_mu
                x, pj, qj = self._xyz()
                res1 = self._sums(x) - self.b - self.a * self.z - self.y + self.mu
                res2 = self.mu * self.lam - epsi
                err = float(max(np.max(np.abs(res1)), np.max(np.abs(res2))))
            epsi *= 0.1
        return x

    @staticmethod
    def design_change(x, xold):
        """MMA.cc:407-426."""
        return float(np.max(np.abs(x - xold)))


# --------------------------------------------------------------------------
# whole design loop (main.cc:48-123), small grids only
# --------------------------------------------------------------------------

def run_design_loop(g: Grid, n_iter, volfrac=0.12, penal=3.0, Emin=1e-9, Emax=1.0, rmin=0.08, filter_type=1,
                    Xmin=0.0, Xmax=1.0, movlim=0.2, projection=False, beta=0.1, beta_final=48.0, eta=0.0,
                    solve_tol=1e-10, log=None):
    """Restatement of main.cc:48-123 with tight solves; returns per-iteration history."""
    KE = hex8_ke(g.dx, g.dy, g.dz, g.nu)
    N, RHS = cantilever_bc_load(g)
    flt = FilterOracle(g, filter_type, rmin)
    if filter_type == 0:
        Xmin = 0.001  # TopOpt.cc:357-359
    x = np.full(g.nel, volfrac)
    xold = x.copy()
    mma = MMAOracle(g.nel, 1)
    xt, xp = flt.filter_project(x, projection, beta, eta)
    U = np.zeros(g.ndof)
    fscale = None
    hist = []
    ch = 1.0
    for itr in range(1, n_iter + 1):
        K = assemble_K(g, KE, xp, Emin, Emax, penal, N)
        U, its, info = solve_state(K, N * RHS, tol=solve_tol, x0=U)
        fx, gx, dfdx, dgdx = objective_sens(g, KE, U, xp, Emin, Emax, penal, volfrac)
        if itr == 1:
            fscale = 10.0 / fx
        dfdx = dfdx * fscale
        dfdx, dg = flt.gradients(x, xt, dfdx, [dgdx], projection, beta, eta)
        xmin, xmax = mma.set_outer_movelimit(Xmin, Xmax, movlim, x)
        xnew = mma.update(x, dfdx, gx, dg, xmin, xmax)
        ch = mma.design_change(xnew, xold)
        xold = xnew.copy()
        x = xnew
        if projection:
            beta, _ = increase_beta(beta, beta_final, gx, itr, ch)
        xt, xp = flt.filter_project(x, projection, beta, eta)
        rec = dict(itr=itr, fx=fx, fscaled=fx * fscale, gx=gx, ch=ch, mnd=mnd(xp), cg_its=its)
        hist.append(rec)
        if log:
            log(rec)
    return dict(hist=hist, x=x, xTilde=xt, xPhys=xp, U=U, fscale=fscale)
