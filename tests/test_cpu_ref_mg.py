"""The reference's solver configuration restated on the CPU (oracle/c/cpu_ref.c: FGMRES + PCMG with Galerkin coarse
operators + GMRES/SOR smoothers, LinearElasticity.cc:613-746) -- checked against scipy: the Galerkin products against
P^T A P with the oracle's Q1 prolongation, the solve against a direct solve, the cone filter against the explicit H.
This is the code bench.py's reference arm times; it is never part of the product path."""
import numpy as np
import pytest
import scipy.sparse.linalg as spl


@pytest.fixture(scope="module")
def setup(oracle):
    import cpu_ref
    shape = (17, 9, 9)
    g = oracle.Grid(*shape)
    cr = cpu_ref.CpuRef(*shape, nthreads=2)
    x = np.random.default_rng(0).uniform(0.05, 1.0, g.nel)
    cr.assemble(x)
    cr.mg_setup(3)
    yield g, cr, x
    cr.close()


def test_galerkin_products_match_ptap(oracle, setup):
    g, cr, x = setup
    A0 = cr.level_matrix(0)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, _ = oracle.cantilever_bc_load(g)
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
    assert abs(A0 - K).max() <= 1e-14 * abs(K).max()
    P1, P2 = oracle.q1_prolongation(9, 5, 5), oracle.q1_prolongation(5, 3, 3)
    A1 = (P1.T @ A0 @ P1).tocsr()
    A2 = (P2.T @ A1 @ P2).tocsr()
    assert abs(A1 - cr.level_matrix(1)).max() <= 1e-14 * abs(A1).max()
    assert abs(A2 - cr.level_matrix(2)).max() <= 1e-14 * abs(A2).max()


def test_fgmres_mg_solves_the_state_problem(oracle, setup):
    g, cr, x = setup
    N, rhs = oracle.cantilever_bc_load(g)
    K = cr.level_matrix(0)
    Ue = spl.spsolve(K.tocsc(), N * rhs)
    U, its, rel, _ = cr.fgmres_mg(rtol=1e-10)
    assert rel <= 1e-10 and 0 < its <= 40
    assert np.abs(U - Ue).max() <= 1e-8 * np.abs(Ue).max()
    # the reference's tolerance: a handful of iterations, and the warm start (KSPSetInitialGuessNonzero) needs none
    U5, its5, rel5, _ = cr.fgmres_mg(rtol=1e-5)
    assert rel5 <= 1e-5 and its5 <= its
    _, its0, _, _ = cr.fgmres_mg(U=U, rtol=1e-5)
    assert its0 == 0


def test_mesh_not_compatible_with_levels():
    import cpu_ref
    cr = cpu_ref.CpuRef(11, 7, 7, nthreads=1)  # 10 x 6 x 6 elements: two coarsenings do not divide
    cr.assemble(np.full(cr.nel, 0.5))
    with pytest.raises(ValueError):
        cr.mg_setup(3)
    cr.close()


@pytest.mark.parametrize("rmin", [0.13, 0.3])
def test_cone_filter_matches_explicit_h(oracle, setup, rmin):
    g, cr, x = setup
    H, Hs, _ = oracle.filter_matrix(g, rmin)
    assert np.abs(cr.cone_filter(rmin, x, 0)[0] - (H @ x) / Hs).max() <= 1e-13
    assert np.abs(cr.cone_filter(rmin, x, 1)[0] - H @ (x / Hs)).max() <= 1e-13
