"""CPU tests: the two independent oracles (numpy restatement and C/OpenMP restatement) agree with
each other and with the committed known answers (tests/golden/appendix_c.json)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "appendix_c.json")))


@pytest.fixture(scope="module")
def cpu_ref():
    so = os.path.join(ROOT, "oracle", "c", "libcpu_ref.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "c/libcpu_ref.so"], check=True)
    import cpu_ref
    return cpu_ref


def test_ke_golden(oracle):
    ke = oracle.hex8_ke(1, 1, 1, 0.3)
    assert np.allclose(ke[0, :6], GOLD["KE_unit_nu03_row0"], rtol=0, atol=1e-15)
    assert np.abs(ke - ke.T).max() < 1e-15 and np.abs(ke.sum(1)).max() < 1e-14
    kf = oracle.pde_filter_ke(1 / 32, 1 / 32, 1 / 32, 0.08 / 2 / np.sqrt(3))
    assert np.allclose(kf[0], GOLD["KF_default_row0"], rtol=1e-14)
    assert np.allclose(kf.sum(1), (1 / 32) ** 3 / 8)


def test_pde_matrix_is_quadrature(oracle):
    """The closed form of PDEFilter.cc:472-565 equals R^2 int gradN.gradN + int N N by 2x2x2 Gauss quadrature."""
    dx, dy, dz, R = 0.3, 0.2, 0.5, 0.17
    gp = 1 / np.sqrt(3)
    sg = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)
    K = np.zeros((8, 8))
    for xi in (-gp, gp):
        for et in (-gp, gp):
            for ze in (-gp, gp):
                N = 0.125 * (1 + sg[:, 0] * xi) * (1 + sg[:, 1] * et) * (1 + sg[:, 2] * ze)
                G = np.stack([0.125 * sg[:, 0] * (1 + sg[:, 1] * et) * (1 + sg[:, 2] * ze) * 2 / dx,
                              0.125 * sg[:, 1] * (1 + sg[:, 0] * xi) * (1 + sg[:, 2] * ze) * 2 / dy,
                              0.125 * sg[:, 2] * (1 + sg[:, 0] * xi) * (1 + sg[:, 1] * et) * 2 / dz])
                K += (dx * dy * dz / 8) * (R * R * G.T @ G + np.outer(N, N))
    assert np.abs(K - oracle.pde_filter_ke(dx, dy, dz, R)).max() < 1e-15


@pytest.mark.parametrize("shape,xc", [((9, 5, 5), (0, 2, 0, 1, 0, 1)), ((13, 8, 6), (0, 1.3, 0, 0.7, 0, 0.45))])
def test_c_port_vs_numpy_oracle(oracle, cpu_ref, shape, xc):
    g = oracle.Grid(*shape, xc=xc)
    cr = cpu_ref.CpuRef(*shape, xc=xc)
    assert cr.ndof == g.ndof and cr.nel == g.nel
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    assert np.abs(cr.ke() - KE).max() <= 1e-15 * np.abs(KE).max() * 24
    N, RHS = oracle.cantilever_bc_load(g)
    assert np.array_equal(cr.rhs(), RHS)
    rng = np.random.default_rng(0)
    x = rng.uniform(0.05, 1, g.nel)
    cr.assemble(x)
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
    assert cr.nnz >= K.nnz
    u = rng.uniform(-1, 1, g.ndof)
    y, _ = cr.spmv(u)
    assert np.abs(y - K @ u).max() <= 1e-13 * np.abs(K @ u).max()
    U, its, rr, _ = cr.pcg(rtol=1e-12)
    U0, _, _ = oracle.solve_state(K, N * RHS, tol=1e-13)
    assert np.abs(U - U0).max() <= 1e-8 * np.abs(U0).max()
    fx, gx, df, _ = cr.objective(U, x)
    fx0, gx0, df0, _ = oracle.objective_sens(g, KE, U0, x, 1e-9, 1.0, 3.0, 0.12)
    assert abs(fx - fx0) <= 1e-9 * abs(fx0) and abs(gx - gx0) < 1e-14
    assert np.abs(df - df0).max() <= 1e-8 * np.abs(df0).max()
    cr.close()


def test_c_port_golden_compliance(cpu_ref):
    cr = cpu_ref.CpuRef(49, 25, 25)
    x = np.full(cr.nel, 0.12)
    cr.assemble(x)
    U, its, rr, _ = cr.pcg(rtol=1e-11)
    fx, gx, df, _ = cr.objective(U, x)
    assert abs(fx - GOLD["48x24x24"]["fx"]) <= 1e-9 * fx
    assert abs(df.sum() - GOLD["48x24x24"]["sum_dfdx"]) <= 1e-8 * abs(df.sum())
    cr.close()


def test_filter_oracle_properties(oracle):
    g = oracle.Grid(13, 9, 7)
    H, Hs, s = oracle.filter_matrix(g, 0.45)
    assert (H - H.T).nnz == 0 or abs(H - H.T).max() < 1e-15
    assert np.allclose(H @ np.ones(g.nel), Hs)
    assert s == oracle.filter_elemconn(g, 0.45)
    x = np.linspace(0, 1, 11)
    assert np.allclose(oracle.smooth_projection(x, 1e-9, 0.3), x, atol=1e-8)  # beta -> 0 is the identity
    d = (oracle.smooth_projection(x + 1e-6, 4.0, 0.3) - oracle.smooth_projection(x - 1e-6, 4.0, 0.3)) / 2e-6
    assert np.allclose(d, oracle.chainrule_smooth_projection(x, 4.0, 0.3), rtol=1e-6)


def test_q1_prolongation_partition_of_unity(oracle):
    P = oracle.q1_prolongation(3, 2, 2, ndof=1)
    assert P.shape == (5 * 3 * 3, 3 * 2 * 2)
    assert np.allclose(P @ np.ones(12), 1.0)
