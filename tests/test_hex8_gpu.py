"""GPU parity of the matrix-free hex8 operator, the PCG state solve, objective and sensitivities
against the CPU oracle (oracle/topopt_oracle.py).  All calls go through the C-ABI (ctypes)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_c.json")))


def densities(kind, g, rng):
    if kind == "uniform":
        return np.full(g.nel, 0.12)
    if kind == "random":
        return rng.uniform(0.0, 1.0, g.nel)
    # checker of 4^3 blocks, high contrast
    i, j, k = np.meshgrid(np.arange(g.nex), np.arange(g.ney), np.arange(g.nez), indexing="ij")
    chk = (((i // 4) + (j // 4) + (k // 4)) % 2).astype(float)
    return chk.transpose(2, 1, 0).ravel()


@pytest.mark.parametrize("shape,xc", [((9, 5, 5), (0, 2, 0, 1, 0, 1)), ((34, 21, 13), (0, 2, 0, 1, 0, 1)),
                                      ((65, 33, 33), (0, 2, 0, 1, 0, 1)), ((20, 12, 9), (0, 1.3, 0, 0.7, 0, 0.45))])
def test_ke_matches_oracle(oracle, shape, xc):
    from topopt_in_petsc_b200 import Context, LinearElasticity
    g = oracle.Grid(*shape, xc=xc)
    ctx = Context(*shape, xc=xc)
    phys = LinearElasticity(ctx, nu=0.3)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    assert np.abs(phys.GetKE() - KE).max() <= 1e-15 * np.abs(KE).max() * 24
    phys.close(); ctx.close()


@pytest.mark.parametrize("shape,xc", [((9, 5, 5), (0, 2, 0, 1, 0, 1)), ((34, 21, 13), (0, 2, 0, 1, 0, 1)),
                                      ((65, 33, 33), (0, 2, 0, 1, 0, 1)), ((20, 12, 9), (0, 1.3, 0, 0.7, 0, 0.45))])
@pytest.mark.parametrize("dens", ["uniform", "random", "checker"])
def test_spmv_parity(oracle, shape, xc, dens):
    """y = N.(K(N.u)) + (1-N).u  vs the oracle; every SpMV kernel; tolerance 1e-12 relative (fp64)."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    rng = np.random.default_rng(1234)
    g = oracle.Grid(*shape, xc=xc)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, _ = oracle.cantilever_bc_load(g)
    x = densities(dens, g, rng)
    E = oracle.simp_modulus(x, 1e-9, 1.0, 3.0)
    u = rng.uniform(-1, 1, g.ndof)
    yref = oracle.spmv_matrix_free(g, KE, E, N, u)
    ctx = Context(*shape, xc=xc)
    assert ctx.nel == g.nel and ctx.ndof == g.ndof  # bit-exact counts (TopOpt.cc:172-175)
    scale = np.abs(yref).max()
    for kernel in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11):
        phys = LinearElasticity(ctx, nu=0.3, opts=solver_opts(spmv_kernel=kernel))
        phys.SetDensity(x, 1e-9, 1.0, 3.0)
        y = phys.SpMV(u)
        assert np.abs(y - yref).max() <= 1e-12 * scale, (kernel, np.abs(y - yref).max() / scale)
        d = phys.Diag()
        K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
        assert np.abs(d - K.diagonal()).max() <= 1e-13 * np.abs(K.diagonal()).max()
        phys.close()
    ctx.close()


def test_spmv_linearity_large():
    """Size-independent property at a larger grid: A(au+bv) = aAu + bAv, and A symmetric: v.Au = u.Av."""
    from topopt_in_petsc_b200 import Context, LinearElasticity
    rng = np.random.default_rng(7)
    shape = (129, 65, 65)
    ctx = Context(*shape)
    phys = LinearElasticity(ctx)
    x = rng.uniform(0, 1, ctx.nel)
    phys.SetDensity(x, 1e-9, 1.0, 3.0)
    u, v = rng.uniform(-1, 1, ctx.ndof), rng.uniform(-1, 1, ctx.ndof)
    Au, Av = phys.SpMV(u), phys.SpMV(v)
    Aw = phys.SpMV(0.3 * u - 1.7 * v)
    s = np.abs(Au).max()
    assert np.abs(Aw - (0.3 * Au - 1.7 * Av)).max() <= 1e-12 * s
    assert abs(v @ Au - u @ Av) <= 1e-11 * abs(v @ Au)
    phys.close(); ctx.close()


@pytest.mark.parametrize("shape,key", [((49, 25, 25), "48x24x24"), ((65, 33, 33), "64x32x32")])
def test_iteration1_compliance_golden(shape, key):
    """Iteration-1 compliance of the reference cantilever at uniform x=0.12.
    Solver tolerance 1e-10 -> fx within 1e-8 relative (fx error is second order in the residual)."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    ctx = Context(*shape)
    phys = LinearElasticity(ctx, opts=solver_opts(rtol=1e-10, maxits=20000))
    x = np.full(ctx.nel, 0.12)
    dfdx, dgdx = np.empty(ctx.nel), np.empty(ctx.nel)
    fx, gx = phys.ComputeObjectiveConstraintsSensitivities(dfdx, dgdx, x, 1e-9, 1.0, 3.0, 0.12)
    gold = GOLD[key]
    assert abs(fx - gold["fx"]) <= 1e-8 * gold["fx"]
    assert abs(gx) < 1e-12
    assert abs(dfdx.sum() - gold["sum_dfdx"]) <= 1e-7 * abs(gold["sum_dfdx"])
    assert abs(dfdx.min() - gold["min_dfdx"]) <= 1e-6 * abs(gold["min_dfdx"])
    assert np.allclose(dgdx, 1.0 / ctx.nel)
    rhs = phys.GetRHS()
    assert abs(np.linalg.norm(rhs) - gold["rhs_norm"]) <= 1e-12 and abs(rhs.sum() - gold["rhs_sum"]) <= 1e-15
    U = phys.GetStateField()
    nxn = shape[0]
    assert abs(U[3 * (nxn - 1) + 2] - gold["uz_tip"]) <= 1e-6 * abs(gold["uz_tip"])
    st = phys.stats()
    assert st.rel_residual <= 1e-10 and st.iterations > 0
    phys.close(); ctx.close()


@pytest.mark.parametrize("dens", ["random", "checker"])
def test_solve_obj_sens_vs_oracle(oracle, dens):
    """State solve + objective + sensitivities on a non-uniform design vs the oracle (tight solves)."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    rng = np.random.default_rng(3)
    shape = (25, 13, 13)
    g = oracle.Grid(*shape)
    x = densities(dens, g, rng)
    if dens == "random":
        x = 0.05 + 0.9 * x
    else:
        x = 0.001 + 0.998 * x
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, RHS = oracle.cantilever_bc_load(g)
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
    import scipy.sparse.linalg as spla
    U = spla.spsolve(K.tocsc(), N * RHS)
    fx0, gx0, df0, dg0 = oracle.objective_sens(g, KE, U, x, 1e-9, 1.0, 3.0, 0.12)
    ctx = Context(*shape)
    phys = LinearElasticity(ctx, opts=solver_opts(rtol=1e-11, maxits=50000))
    dfdx, dgdx = np.empty(ctx.nel), np.empty(ctx.nel)
    fx, gx = phys.ComputeObjectiveConstraintsSensitivities(dfdx, dgdx, x, 1e-9, 1.0, 3.0, 0.12)
    assert abs(fx - fx0) <= 1e-7 * abs(fx0)
    assert abs(gx - gx0) <= 1e-13
    assert np.abs(dfdx - df0).max() <= 1e-6 * np.abs(df0).max()
    Ug = phys.GetStateField()
    assert np.abs(Ug - U).max() <= 1e-6 * np.abs(U).max()
    # warm start: a second solve from the converged state needs (almost) no iterations
    st = phys.SolveState(x, 1e-9, 1.0, 3.0)
    assert st.iterations <= 1
    phys.close(); ctx.close()


def test_custom_bc_load(oracle):
    """Generalised BC/load entry point: clamp the z=0 plane, point load at a top corner."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    import scipy.sparse.linalg as spla
    shape = (13, 9, 11)
    g = oracle.Grid(*shape)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N = np.ones(g.ndof)
    N[:3 * g.nxn * g.nyn] = 0.0
    RHS = np.zeros(g.ndof)
    RHS[3 * g.node_id(g.nxn - 1, g.nyn - 1, g.nzn - 1) + 1] = 0.5
    x = np.full(g.nel, 0.4)
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
    U = spla.spsolve(K.tocsc(), N * RHS)
    ctx = Context(*shape)
    phys = LinearElasticity(ctx, opts=solver_opts(rtol=1e-12, maxits=50000))
    phys.SetBCLoad((1 - N).astype(np.uint8), RHS)
    phys.SolveState(x, 1e-9, 1.0, 3.0)
    assert np.abs(phys.GetStateField() - U).max() <= 1e-7 * np.abs(U).max()
    phys.close(); ctx.close()


@pytest.mark.parametrize("dens", ["uniform", "random", "checker"])
def test_gmg_pcg_matches_direct(oracle, dens):
    """PCG preconditioned by the geometric-multigrid V-cycle converges to the same state as a direct solve,
    in far fewer iterations than Jacobi-PCG."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(3)
    shape = (33, 17, 17)
    g = oracle.Grid(*shape)
    x = densities(dens, g, rng)
    x = 0.05 + 0.9 * x if dens == "random" else (0.001 + 0.998 * x if dens == "checker" else x)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, RHS = oracle.cantilever_bc_load(g)
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
    U = spla.spsolve(K.tocsc(), N * RHS)
    ctx = Context(*shape)
    jac = LinearElasticity(ctx, opts=solver_opts(pc=0, rtol=1e-8, maxits=50000))
    sj = jac.SolveState(x, 1e-9, 1.0, 3.0)
    mg = LinearElasticity(ctx, opts=solver_opts(pc=1, nlvls=3, rtol=1e-8, maxits=500))
    sm = mg.SolveState(x, 1e-9, 1.0, 3.0)
    print(f"[{dens}] jacobi its={sj.iterations} gmg its={sm.iterations}")
    assert sm.rel_residual <= 1e-8 and sm.iterations < 500
    assert sm.iterations * 5 < sj.iterations
    Ug = mg.GetStateField()
    assert np.abs(Ug - U).max() <= 1e-5 * np.abs(U).max()
    jac.close(); mg.close(); ctx.close()


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 2, 4), (5, 3, 2), (2, 7, 3), (31, 2, 2)])
def test_tiny_and_ragged_grids(oracle, shape):
    """Edge cases: single-element directions, grids smaller than one tile, odd sizes."""
    from topopt_in_petsc_b200 import Context, Filter, LinearElasticity, solver_opts
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(9)
    g = oracle.Grid(*shape)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, RHS = oracle.cantilever_bc_load(g)
    x = rng.uniform(0.2, 1.0, g.nel)
    E = oracle.simp_modulus(x, 1e-9, 1.0, 3.0)
    u = rng.uniform(-1, 1, g.ndof)
    yref = oracle.spmv_matrix_free(g, KE, E, N, u)
    ctx = Context(*shape)
    assert (ctx.nel, ctx.ndof) == (g.nel, g.ndof)
    for kernel in (1, 5, 9):
        phys = LinearElasticity(ctx, opts=solver_opts(spmv_kernel=kernel, rtol=1e-12, maxits=5000))
        phys.SetDensity(x, 1e-9, 1.0, 3.0)
        assert np.abs(phys.SpMV(u) - yref).max() <= 1e-12 * np.abs(yref).max()
        phys.close()
    phys = LinearElasticity(ctx, opts=solver_opts(rtol=1e-12, maxits=5000))
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
    U = spla.spsolve(K.tocsc(), N * RHS)
    df, dg = np.empty(g.nel), np.empty(g.nel)
    fx, gx = phys.ComputeObjectiveConstraintsSensitivities(df, dg, x, 1e-9, 1.0, 3.0, 0.12)
    fx0, gx0, df0, _ = oracle.objective_sens(g, KE, U, x, 1e-9, 1.0, 3.0, 0.12)
    assert abs(fx - fx0) <= 1e-8 * abs(fx0) and np.abs(df - df0).max() <= 1e-7 * np.abs(df0).max()
    fo = oracle.FilterOracle(g, 1, 0.6)
    f = Filter(ctx, 1, 0.6)
    xt, xp = np.empty(g.nel), np.empty(g.nel)
    f.FilterProject(x, xt, xp, False, 1.0, 0.0)
    assert np.abs(xt - fo.filter_project(x, False, 1.0, 0.0)[0]).max() <= 1e-13
    f.close(); phys.close(); ctx.close()


def test_error_paths():
    """Error convention: non-zero return + message, no crash (the reference returns PetscErrorCode)."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, Filter, TobError, solver_opts
    with pytest.raises(TobError):
        Context(1, 5, 5)
    ctx = Context(9, 5, 5)
    phys = LinearElasticity(ctx)
    with pytest.raises(TobError):
        phys.SpMV(np.zeros(ctx.ndof))          # density not set yet
    with pytest.raises(TobError):
        LinearElasticity(ctx, opts=solver_opts(pc=1, nlvls=4))  # 8x4x4 elements cannot be coarsened 3 times
    f = Filter(ctx, 7, 0.1)                    # unknown filter type = no filtering (Filter.cc:104-107)
    x = np.linspace(0, 1, ctx.nel); xt, xp = np.empty(ctx.nel), np.empty(ctx.nel)
    f.FilterProject(x, xt, xp, False, 1.0, 0.0)
    assert np.array_equal(xt, x) and np.array_equal(xp, x)
    with pytest.raises(TobError):
        f.GetHs()
    f.close(); phys.close(); ctx.close()


# ---- BASELINE configs[1] size: the default kernel's z-chunking, wave logic and (in the solver) L2 hints engage here -------
def _spmv_matrix_free_fast(oracle, g, KE, E, N, u):
    """oracle.spmv_matrix_free with the scatter done by bincount (np.add.at takes minutes at 1.8 M elements)."""
    ed = g.edof()
    ye = ((N * u)[ed] @ KE.T) * E[:, None]
    y = np.bincount(ed.ravel(), weights=ye.ravel(), minlength=g.ndof)
    return N * y + (1.0 - N) * u


@pytest.mark.parametrize("dens", ["uniform", "random", "checker"])
def test_spmv_parity_cfg2_size(oracle, dens):
    """193 x 97 x 97 nodes (BASELINE configs[1]): default SpMV kernel and diag(K) vs the numpy oracle, <= 1e-12 relative."""
    from topopt_in_petsc_b200 import Context, LinearElasticity
    rng = np.random.default_rng(1234)
    shape = (193, 97, 97)
    g = oracle.Grid(*shape)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, _ = oracle.cantilever_bc_load(g)
    x = densities(dens, g, rng)
    E = oracle.simp_modulus(x, 1e-9, 1.0, 3.0)
    u = rng.uniform(-1, 1, g.ndof)
    # small-grid check of the fast scatter against the oracle's own function
    gs = oracle.Grid(9, 5, 5)
    us, Es = rng.uniform(-1, 1, gs.ndof), rng.uniform(0, 1, gs.nel)
    Ns, _ = oracle.cantilever_bc_load(gs)
    KEs = oracle.hex8_ke(gs.dx, gs.dy, gs.dz, 0.3)
    assert np.abs(_spmv_matrix_free_fast(oracle, gs, KEs, Es, Ns, us) - oracle.spmv_matrix_free(gs, KEs, Es, Ns, us)).max() < 1e-13
    yref = _spmv_matrix_free_fast(oracle, g, KE, E, N, u)
    ctx = Context(*shape)
    assert ctx.nel == g.nel == 1769472 and ctx.ndof == g.ndof == 5447811
    phys = LinearElasticity(ctx, nu=0.3)
    phys.SetDensity(x, 1e-9, 1.0, 3.0)
    y = phys.SpMV(u)
    assert np.abs(y - yref).max() <= 1e-12 * np.abs(yref).max()
    # diag(K) = KE_cc * sum of the 8 adjacent moduli on free dofs, 1 on clamped ones (LinearElasticity.cc:510-538)
    ed = g.edof()
    dref = np.bincount(ed.ravel(), weights=(E[:, None] * np.diag(KE)[None, :]).ravel(), minlength=g.ndof)
    dref = N * dref + (1.0 - N)
    assert np.abs(phys.Diag() - dref).max() <= 1e-13 * dref.max()
    phys.close(); ctx.close()


def test_iteration1_compliance_cfg2_vs_cpu_reference_solver():
    """Iteration-1 compliance of BASELINE configs[1] (192x96x96, uniform 0.12): the product's Jacobi-PCG (default kernel,
    L2 hints and CUDA-graphed iterations on) and its multigrid-PCG, both to 1e-10, against the C restatement of the
    reference's FGMRES + Galerkin-PCMG solver on the assembled CSR matrix (oracle/c/cpu_ref.c), also to 1e-10: FGMRES
    minimises the residual, not the energy error, so the compliance f.u is only FIRST order in it (measured: rtol 1e-8 leaves
    1.4e-8 .. 5e-8 relative in fx depending on the thread count, rtol 1e-10 leaves 7e-11).  Checked to 1e-8 relative."""
    import cpu_ref
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    shape = (193, 97, 97)
    ctx = Context(*shape)
    x = np.full(ctx.nel, 0.12)
    dfdx, dgdx = np.empty(ctx.nel), np.empty(ctx.nel)
    res = {}
    for name, opts in (("jacobi", solver_opts(pc=0, rtol=1e-10, maxits=200000, check_every=50)),
                       ("gmg", solver_opts(pc=1, nlvls=6, rtol=1e-10, maxits=500, cheb_degree=3, coarse_its=20))):
        phys = LinearElasticity(ctx, opts=opts)
        fx, gx = phys.ComputeObjectiveConstraintsSensitivities(dfdx, dgdx, x, 1e-9, 1.0, 3.0, 0.12)
        st = phys.stats()
        assert st.rel_residual <= 1e-10
        res[name] = (fx, dfdx.copy(), st.iterations)
        phys.close()
    ctx.close()
    # physical cores, at most 32: the C restatement is memory-bound (8 cores: 50 s for this solve) and its barriers spin -- with one
    # thread per *logical* CPU of a 64-core box the same solve took ten minutes
    try:
        import psutil
        ncores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except ImportError:
        ncores = max(1, (os.cpu_count() or 2) // 2)
    cr = cpu_ref.CpuRef(*shape, nthreads=min(32, ncores))
    cr.assemble(x)
    cr.mg_setup(4)
    U, its, rel, _ = cr.fgmres_mg(rtol=1e-10, maxits=200)  # 13-15 iterations, about a minute on 8 cores
    assert rel <= 1e-10
    fx0, gx0, df0, _ = cr.objective(U, x)
    cr.close()
    print(f"cfg2 iteration 1: cpu fx={fx0:.12g} ({its} FGMRES its), jacobi {res['jacobi'][0]:.12g} ({res['jacobi'][2]} its), "
          f"gmg {res['gmg'][0]:.12g} ({res['gmg'][2]} its)")
    for name in res:
        assert abs(res[name][0] - fx0) <= 1e-8 * fx0, name
        assert np.abs(res[name][1] - df0).max() <= 1e-6 * np.abs(df0).max(), name


def test_split_objective_and_sensitivity_entry_points(oracle):
    """ComputeObjectiveConstraints (LinearElasticity.cc:225-297) and ComputeSensitivities (:299-361) through
    tob_elast_obj / tob_elast_sens give what the fused call gives, and what the oracle gives."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(11)
    shape = (21, 11, 9)
    g = oracle.Grid(*shape)
    x = 0.05 + 0.9 * rng.uniform(0, 1, g.nel)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, RHS = oracle.cantilever_bc_load(g)
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
    U = spla.spsolve(K.tocsc(), N * RHS)
    fx0, gx0, df0, dg0 = oracle.objective_sens(g, KE, U, x, 1e-9, 1.0, 3.0, 0.3)
    ctx = Context(*shape)
    phys = LinearElasticity(ctx, opts=solver_opts(rtol=1e-12, maxits=50000))
    fx, gx = phys.ComputeObjectiveConstraints(x, 1e-9, 1.0, 3.0, 0.3)
    assert abs(fx - fx0) <= 1e-8 * fx0 and abs(gx - gx0) <= 1e-13
    dfdx, dgdx = np.empty(g.nel), np.empty(g.nel)
    phys.ComputeSensitivities(dfdx, dgdx, x, 1e-9, 1.0, 3.0, 0.3)  # reuses the state of the solve above (:317-318)
    assert np.abs(dfdx - df0).max() <= 1e-7 * np.abs(df0).max()
    assert np.allclose(dgdx, dg0, rtol=0, atol=1e-18)
    d2, g2 = np.empty(g.nel), np.empty(g.nel)
    fx2, gx2 = phys.ComputeObjectiveConstraintsSensitivities(d2, g2, x, 1e-9, 1.0, 3.0, 0.3)
    assert abs(fx2 - fx) <= 1e-12 * fx and gx2 == gx and np.abs(d2 - dfdx).max() <= 1e-12 * np.abs(dfdx).max()
    phys.close(); ctx.close()


@pytest.mark.parametrize("dens", ["uniform", "random"])
def test_single_reduction_pcg_on_one_rank(oracle, dens, monkeypatch):
    """The Chronopoulos-Gear Jacobi-PCG (cgsr.cuh: 2 launches per iteration, sums published by the SpMV's last CTA), forced
    on one rank with TOB_PCG_SR=1, converges to the same state as a direct solve; iteration count close to classical PCG."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(5)
    shape = (33, 17, 17)
    g = oracle.Grid(*shape)
    x = densities(dens, g, rng)
    x = 0.05 + 0.9 * x if dens == "random" else x
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, RHS = oracle.cantilever_bc_load(g)
    U = spla.spsolve(oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N).tocsc(), N * RHS)
    ctx = Context(*shape)
    monkeypatch.setenv("TOB_PCG_SR", "0")
    cl = LinearElasticity(ctx, opts=solver_opts(rtol=1e-8, maxits=50000, check_every=10))
    s0 = cl.SolveState(x, 1e-9, 1.0, 3.0)
    monkeypatch.setenv("TOB_PCG_SR", "1")
    sr = LinearElasticity(ctx, opts=solver_opts(rtol=1e-8, maxits=50000, check_every=10))
    s1 = sr.SolveState(x, 1e-9, 1.0, 3.0)
    print(f"[{dens}] classical its={s0.iterations} single-reduction its={s1.iterations}")
    assert s1.rel_residual <= 1e-8
    assert s1.iterations <= 1.15 * s0.iterations + 24
    assert np.abs(sr.GetStateField() - U).max() <= 1e-5 * np.abs(U).max()
    s2 = sr.SolveState(0.9 * x + 0.05, 1e-9, 1.0, 3.0)  # warm-started second solve through the captured graph
    assert s2.rel_residual <= 1e-8
    cl.close(); sr.close(); ctx.close()


def test_galerkin_coarse_operators_match_ptap(oracle, monkeypatch):
    """Galerkin levels of the multigrid hierarchy (galerkin.cu; the reference: PCMGSetGalerkin, LinearElasticity.cc:702):
    with TOB_GMG_GALERKIN_FROM=1 level 1 is assembled from the matrix-free fine level, level 2 from level 1's stencil.
    Both must equal N (P^T A P) N + (I - N) formed with scipy from the oracle's assembled matrix and Q1 prolongation."""
    import scipy.sparse as sp
    from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
    monkeypatch.setenv("TOB_GMG_GALERKIN_FROM", "1")
    monkeypatch.setenv("TOB_GMG_STENCIL_FP64", "1")  # the coefficients are stored in fp32 by default (a preconditioner); exact parity in fp64
    rng = np.random.default_rng(21)
    shape = (17, 9, 13)
    g = oracle.Grid(*shape)
    x = rng.uniform(0.0, 1.0, g.nel) ** 3
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N0, _ = oracle.cantilever_bc_load(g)
    K0 = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, np.ones(g.ndof))  # unconstrained: the masks enter level by level
    ctx = Context(*shape)
    mg = LinearElasticity(ctx, opts=solver_opts(pc=1, nlvls=3, rtol=1e-8, maxits=200))
    mg.SetDensity(x, 1e-9, 1.0, 3.0)
    assert mg.GmgLevels() == 3 and mg.GmgLevelMatrix(0) is None
    A_prev, dims = K0, (17, 9, 13)
    for lvl in (1, 2):
        nc = tuple((d - 1) // 2 + 1 for d in dims)
        P = oracle.q1_prolongation(*nc)
        gc = oracle.Grid(*nc)
        Nc, _ = oracle.cantilever_bc_load(gc)  # injection of the wall x = xmin
        D = sp.diags(Nc)
        ref = (D @ (P.T @ A_prev @ P) @ D + sp.diags(1.0 - Nc)).tocsr()
        got = mg.GmgLevelMatrix(lvl)
        assert got is not None
        if lvl == 1:
            got1 = got
        assert abs(got - ref).max() <= 1e-12 * abs(ref).max(), (lvl, abs(got - ref).max() / abs(ref).max())
        # the next level is the product of THIS level's operator without its Dirichlet rows: rebuild the unconstrained one
        A_prev, dims = (P.T @ A_prev @ P).tocsr(), nc
    # and the preconditioned solve converges to the direct solution
    import scipy.sparse.linalg as spla
    _, RHS = oracle.cantilever_bc_load(g)
    K = oracle.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N0)
    U = spla.spsolve(K.tocsc(), N0 * RHS)
    st = mg.SolveState(x, 1e-9, 1.0, 3.0)
    assert st.rel_residual <= 1e-8 and st.iterations < 200, st.iterations  # a random cubed field on 16x8x12 elements is a hard, tiny case
    assert np.abs(mg.GetStateField() - U).max() <= 1e-5 * np.abs(U).max()
    mg.close()
    # default storage (fp32 coefficients): the same operators to single precision, the same solution
    monkeypatch.setenv("TOB_GMG_STENCIL_FP64", "0")
    mg32 = LinearElasticity(ctx, opts=solver_opts(pc=1, nlvls=3, rtol=1e-8, maxits=200))
    mg32.SetDensity(x, 1e-9, 1.0, 3.0)
    got32 = mg32.GmgLevelMatrix(1)
    assert abs(got32 - got1).max() <= 1e-6 * abs(got1).max()
    st32 = mg32.SolveState(x, 1e-9, 1.0, 3.0)
    assert st32.rel_residual <= 1e-8 and st32.iterations <= st.iterations + 5
    assert np.abs(mg32.GetStateField() - U).max() <= 1e-5 * np.abs(U).max()
    mg32.close(); ctx.close()


@pytest.mark.parametrize("shape,xc", [((9, 5, 5), (0, 2, 0, 1, 0, 1)), ((34, 21, 13), (0, 2, 0, 1, 0, 1)), ((65, 33, 33), (0, 2, 0, 1, 0, 1)),
                                      ((20, 12, 9), (0, 1.3, 0, 0.7, 0, 0.45))])
@pytest.mark.parametrize("dens", ["uniform", "random", "checker"])
@pytest.mark.parametrize("gen", [1, 2, 3])
def test_spmv_tf32_tensor_core_operator(oracle, shape, xc, dens, gen, monkeypatch):
    """The single-precision tensor-core operator (hex8_tc32.cu: tcgen05.mma kind::tf32, fp32 accumulation in TMEM) applies
    the same operator as the fp64 kernel up to TF32 input rounding (10-bit mantissa): <= 2e-3 of max|y| against the oracle;
    fixed dofs are passed through exactly (to fp32)."""
    from topopt_in_petsc_b200 import Context, LinearElasticity
    # gen 1: element-row GEMM with gather / shared-memory scatter; gen 2: the tensor core reads the nodal planes directly,
    # one skinny product per element position; gen 3: the same operand view with element rows (wide tiles) + transposed node sums
    monkeypatch.setenv("TOB_TC32_GEN", str(gen))
    rng = np.random.default_rng(77)
    g = oracle.Grid(*shape, xc=xc)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3)
    N, _ = oracle.cantilever_bc_load(g)
    x = densities(dens, g, rng)
    E = oracle.simp_modulus(x, 1e-9, 1.0, 3.0)
    u = rng.uniform(-1, 1, g.ndof)
    yref = oracle.spmv_matrix_free(g, KE, E, N, u)
    ctx = Context(*shape, xc=xc)
    phys = LinearElasticity(ctx, nu=0.3)
    phys.SetDensity(x, 1e-9, 1.0, 3.0)
    y = phys.SpMVtf32(u)
    err = np.abs(y - yref).max() / np.abs(yref).max()
    assert err <= 2e-3, err
    assert err > 1e-9  # really a reduced-precision path, not the fp64 kernel in disguise
    fixed = N == 0
    assert np.abs(y[fixed] - u[fixed]).max() <= 1e-6
    # linear to fp32 round-off
    y2 = phys.SpMVtf32(0.5 * u)
    assert np.abs(y2 - 0.5 * y).max() <= 1e-5 * np.abs(yref).max()
    phys.close(); ctx.close()
