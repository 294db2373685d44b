"""GPU parity of the filter kernels (cone filter, Heaviside, MND) and the device MMA against the
CPU oracle, plus a short whole-loop (main.cc:48-123) trajectory comparison."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_c.json")))


@pytest.mark.parametrize("shape,rmin,key", [((65, 33, 33), 0.08, "64x32x32"), ((49, 25, 25), 0.08, "48x24x24")])
def test_hs_golden(shape, rmin, key):
    from topopt_in_petsc_b200 import Context, Filter
    ctx = Context(*shape)
    f = Filter(ctx, 1, rmin)
    assert f.ElemConn() == GOLD[f"elemconn_{key}_r008"]
    hs = f.GetHs()
    assert abs(hs.max() - GOLD[f"Hs_interior_{key}_r008"]) <= 1e-13
    f.close(); ctx.close()


@pytest.mark.parametrize("shape,xc,rmin", [((17, 9, 9), (0, 2, 0, 1, 0, 1), 0.3), ((21, 13, 11), (0, 2, 0, 1, 0, 1), 0.25),
                                           ((13, 9, 7), (0, 1.0, 0, 0.9, 0, 0.5), 0.21), ((9, 5, 5), (0, 2, 0, 1, 0, 1), 5.0),
                                           ((33, 17, 17), (0, 2, 0, 1, 0, 1), 0.2)])  # the last one: ElemConn 3 (7^3 box, tiled kernel <3>)
@pytest.mark.parametrize("ftype", [0, 1, 3])
@pytest.mark.parametrize("proj", [False, True])
def test_filter_project_gradients(oracle, shape, xc, rmin, ftype, proj):
    from topopt_in_petsc_b200 import Context, Filter
    rng = np.random.default_rng(11)
    g = oracle.Grid(*shape, xc=xc)
    fo = oracle.FilterOracle(g, ftype, rmin)
    ctx = Context(*shape, xc=xc)
    f = Filter(ctx, ftype, rmin)
    if ftype in (0, 1):
        assert f.ElemConn() == fo.elemconn
        assert np.abs(f.GetHs() - fo.Hs).max() <= 1e-13 * fo.Hs.max()
    x = rng.uniform(0.01, 1.0, g.nel)
    beta, eta = 4.0, 0.3
    xt, xp = np.empty(g.nel), np.empty(g.nel)
    f.FilterProject(x, xt, xp, proj, beta, eta)
    xt0, xp0 = fo.filter_project(x, proj, beta, eta)
    assert np.abs(xt - xt0).max() <= 1e-13 and np.abs(xp - xp0).max() <= 1e-13
    assert abs(f.GetMND(xp) - oracle.mnd(xp0)) <= 1e-13
    df, dg = rng.uniform(-1, 0, g.nel), np.full(g.nel, 1.0 / g.nel)
    df0, dg0 = fo.gradients(x, xt0, df, [dg], proj, beta, eta)
    f.Gradients(x, xt, df, 1, [dg], proj, beta, eta)
    assert np.abs(df - df0).max() <= 1e-12 * np.abs(df0).max()
    assert np.abs(dg - dg0[0]).max() <= 1e-12 * np.abs(dg0[0]).max()
    f.close(); ctx.close()


def test_filter_constants_and_adjoint():
    """Size-independent properties at a larger grid: the density filter preserves constants and
    <H(g/Hs), x> = <g, (Hx)/Hs> (Gradients is the adjoint of FilterProject)."""
    from topopt_in_petsc_b200 import Context, Filter
    rng = np.random.default_rng(5)
    shape = (97, 49, 49)
    ctx = Context(*shape)
    f = Filter(ctx, 1, 3.0 / 48)  # 3 cells
    n = ctx.nel
    one = np.full(n, 0.37)
    xt, xp = np.empty(n), np.empty(n)
    f.FilterProject(one, xt, xp, False, 1.0, 0.0)
    assert np.abs(xt - 0.37).max() <= 1e-14 and np.array_equal(xt, xp)
    x, gvec = rng.uniform(0, 1, n), rng.uniform(-1, 1, n)
    f.FilterProject(x, xt, xp, False, 1.0, 0.0)
    gf, dum = gvec.copy(), np.zeros(n)
    f.Gradients(x, xt, gf, 1, [dum], False, 1.0, 0.0)
    assert abs(gf @ x - gvec @ xt) <= 1e-11 * abs(gvec @ xt)
    f.close(); ctx.close()


def test_increase_beta(oracle):
    from topopt_in_petsc_b200 import Context, Filter
    ctx = Context(9, 5, 5)
    f = Filter(ctx, 1, 0.3)
    for beta, bf, gx, itr, ch in [(0.1, 48, -1e-3, 10, 0.2), (6.5, 48, -1e-3, 3, 0.001), (7.0, 48, -1e-3, 20, 0.5),
                                  (45.0, 48, -1.0, 30, 0.5), (3.0, 48, 1e-3, 10, 0.0), (3.0, 48, -1.0, 7, 0.5)]:
        b1, c1 = f.IncreaseBeta(beta, bf, gx, itr, ch)
        b0, c0 = oracle.increase_beta(beta, bf, gx, itr, ch)
        assert b1 == b0 and c1 == c0
    f.close(); ctx.close()


@pytest.mark.parametrize("m", [1, 2])
def test_mma_vs_oracle(oracle, m):
    """Device MMA vs the oracle restatement of MMA.cc over several outer iterations of a synthetic problem."""
    from topopt_in_petsc_b200 import Context, MMA
    rng = np.random.default_rng(21)
    shape = (17, 9, 9)
    ctx = Context(*shape)
    n = ctx.nel
    mo = oracle.MMAOracle(n, m)
    mg = MMA(ctx, m)
    x = np.full(n, 0.3)
    xg = x.copy()
    xold_o, xold_g = x.copy(), x.copy()
    for it in range(6):
        c0 = rng.uniform(0.5, 2.0, n)
        dfdx = -c0 / (x + 0.1) ** 2
        gx = [float(np.mean(x) - 0.3)] + ([float(np.mean(x ** 2) - 0.2)] if m == 2 else [])
        dgdx = [np.full(n, 1.0 / n)] + ([2 * x / n] if m == 2 else [])
        xmin, xmax = mo.set_outer_movelimit(0.0, 1.0, 0.2, x)
        xmin_g, xmax_g = np.empty(n), np.empty(n)
        mg.SetOuterMovelimit(0.0, 1.0, 0.2, xg, xmin_g, xmax_g)
        assert np.array_equal(xmin, xmin_g) and np.array_equal(xmax, xmax_g)
        xnew = mo.update(x, dfdx, np.array(gx), dgdx, xmin, xmax)
        mg.Update(xg, dfdx, np.array(gx), [d.copy() for d in dgdx], xmin_g, xmax_g)
        assert np.abs(xg - xnew).max() <= 1e-9, (it, np.abs(xg - xnew).max())
        ch_o = mo.design_change(xnew, xold_o); xold_o = xnew.copy()
        ch_g = mg.DesignChange(xg, xold_g)
        assert abs(ch_o - ch_g) <= 1e-9 and np.array_equal(xold_g, xg)
        x = xnew.copy(); xg = xnew.copy()  # keep both on the same trajectory
    mg.close(); ctx.close()


def test_design_loop_vs_oracle(oracle):
    """First design iterations of main.cc:48-123 on 16x8x8 with tight solves: compliance history within
    1e-6 relative, design field within 1e-5 (SURVEY 7.3-3 'trajectory' parity level)."""
    from topopt_in_petsc_b200 import Context, LinearElasticity, Filter, MMA, solver_opts
    shape = (17, 9, 9)
    g = oracle.Grid(*shape)
    ref = oracle.run_design_loop(g, 5, rmin=0.3, solve_tol=1e-11)
    ctx = Context(*shape)
    phys = LinearElasticity(ctx, opts=solver_opts(rtol=1e-11, maxits=50000))
    flt = Filter(ctx, 1, 0.3)
    mma = MMA(ctx, 1)
    n = ctx.nel
    x, xt, xp = np.full(n, 0.12), np.empty(n), np.empty(n)
    xold, xmin, xmax = x.copy(), np.empty(n), np.empty(n)
    dfdx, dgdx = np.empty(n), np.empty(n)
    flt.FilterProject(x, xt, xp, False, 0.1, 0.0)
    fscale = None
    for itr in range(1, 6):
        fx, gx = phys.ComputeObjectiveConstraintsSensitivities(dfdx, dgdx, xp, 1e-9, 1.0, 3.0, 0.12)
        if itr == 1:
            fscale = 10.0 / fx
        dfdx *= fscale
        flt.Gradients(x, xt, dfdx, 1, [dgdx], False, 0.1, 0.0)
        mma.SetOuterMovelimit(0.0, 1.0, 0.2, x, xmin, xmax)
        mma.Update(x, dfdx, np.array([gx]), [dgdx], xmin, xmax)
        ch = mma.DesignChange(x, xold)
        flt.FilterProject(x, xt, xp, False, 0.1, 0.0)
        h = ref["hist"][itr - 1]
        assert abs(fx - h["fx"]) <= 1e-6 * h["fx"], (itr, fx, h["fx"])
        assert abs(gx - h["gx"]) <= 1e-9 and abs(ch - h["ch"]) <= 1e-6
    assert np.abs(x - ref["x"]).max() <= 1e-5
    assert np.abs(xp - ref["xPhys"]).max() <= 1e-5
    for o in (mma, flt, phys):
        o.close()
    ctx.close()


@pytest.mark.parametrize("shape,xc,rmin", [((17, 9, 9), (0, 2, 0, 1, 0, 1), 0.3), ((21, 13, 11), (0, 1.0, 0, 0.9, 0, 0.5), 0.21),
                                           ((33, 17, 17), (0, 2, 0, 1, 0, 1), 0.08)])
def test_pde_filter_vs_oracle(oracle, shape, xc, rmin):
    """Helmholtz PDE filter (PDEFilter.cc:189-216): T^T K^-1 (vol T x), solved to rtol 1e-8 on both sides."""
    from topopt_in_petsc_b200 import Context, Filter
    rng = np.random.default_rng(2)
    g = oracle.Grid(*shape, xc=xc)
    po = oracle.PDEFilterOracle(g, rmin)
    ctx = Context(*shape, xc=xc)
    f = Filter(ctx, 2, rmin)
    x = rng.uniform(0, 1, g.nel)
    out = f.PDEApply(x)
    ref = po.apply(x)
    assert np.abs(out - ref).max() <= 1e-7
    st = f.stats()
    assert st.rel_residual <= 1e-8 and 0 < st.iterations < 2000
    # preserves constants (rows of KF sum to vol/8) and is self-adjoint
    assert np.abs(f.PDEApply(np.full(g.nel, 0.4)) - 0.4).max() <= 1e-7
    y = rng.uniform(0, 1, g.nel)
    assert abs(y @ out - x @ f.PDEApply(y)) <= 1e-6 * abs(y @ out)
    # through the Filter surface: clamp to [0,1], projection, gradients
    xt, xp = np.empty(g.nel), np.empty(g.nel)
    f.FilterProject(x, xt, xp, True, 4.0, 0.3)
    fo = oracle.FilterOracle(g, 2, rmin)
    xt0, xp0 = fo.filter_project(x, True, 4.0, 0.3)
    assert np.abs(xt - xt0).max() <= 1e-7 and np.abs(xp - xp0).max() <= 1e-6
    df, dg = rng.uniform(-1, 0, g.nel), np.full(g.nel, 1.0 / g.nel)
    df0, dg0 = fo.gradients(x, xt0, df, [dg], True, 4.0, 0.3)
    f.Gradients(x, xt, df, 1, [dg], True, 4.0, 0.3)
    assert np.abs(df - df0).max() <= 1e-6 * np.abs(df0).max() + 1e-9
    assert np.abs(dg - dg0[0]).max() <= 1e-6 * np.abs(dg0[0]).max() + 1e-12
    f.close(); ctx.close()


def test_pde_filter_multigrid_large_radius(oracle, monkeypatch):
    """rmin = 8 cells on 64x32x32: the V-cycle-preconditioned CG (reference: FGMRES + 3-level PCMG, PDEFilter.cc:276-295,
    327-378) needs a handful of iterations where Jacobi-PCG needs many; both reproduce the oracle's sparse solve."""
    from topopt_in_petsc_b200 import Context, Filter
    rng = np.random.default_rng(4)
    shape = (65, 33, 33)
    g = oracle.Grid(*shape)
    rmin = 8 * g.dx
    x = rng.uniform(0, 1, g.nel)
    ref = oracle.PDEFilterOracle(g, rmin).apply(x)
    ctx = Context(*shape)
    monkeypatch.setenv("TOB_PDE_MG", "0")
    fj = Filter(ctx, 2, rmin)
    oj = fj.PDEApply(x)
    sj = fj.stats()
    monkeypatch.setenv("TOB_PDE_MG", "1")
    fm = Filter(ctx, 2, rmin)
    om = fm.PDEApply(x)
    sm = fm.stats()
    print(f"PDE filter rmin = 8h: Jacobi-PCG {sj.iterations} its {sj.seconds * 1e3:.2f} ms, MG-PCG {sm.iterations} its {sm.seconds * 1e3:.2f} ms")
    assert sm.rel_residual <= 1e-8 and sj.rel_residual <= 1e-8
    assert np.abs(om - ref).max() <= 1e-7 and np.abs(oj - ref).max() <= 1e-7
    assert sm.iterations <= 15 and sm.iterations * 3 < sj.iterations
    assert np.abs(fm.PDEApply(np.full(g.nel, 0.4)) - 0.4).max() <= 1e-7
    fj.close(); fm.close(); ctx.close()
