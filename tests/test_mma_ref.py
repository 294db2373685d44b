"""Pins the MMA restatements against the REFERENCE's own unmodified MMA.cc (compiled from
/root/reference into oracle/_ref/libmma_ref.so by oracle/Makefile):
  - CPU: oracle.MMAOracle == reference MMA over several outer iterations (m = 1, 2)
  - GPU: the device MMA (csrc/mma.cu, through the C-ABI) == reference MMA."""
import numpy as np
import pytest


def problem(n, m, x, rng):
    c0 = rng.uniform(0.5, 2.0, n)
    dfdx = -c0 / (x + 0.1) ** 2
    gx = [float(np.mean(x) - 0.3)] + ([float(np.mean(x ** 2) - 0.2)] if m == 2 else [])
    dgdx = [np.full(n, 1.0 / n)] + ([2 * x / n] if m == 2 else [])
    return dfdx, np.array(gx), dgdx


@pytest.fixture(scope="module")
def mma_ref():
    import mma_ref
    if not mma_ref.available():
        pytest.skip("oracle/_ref/libmma_ref.so not built (needs /root/reference at build time)")
    return mma_ref


@pytest.mark.parametrize("m", [1, 2])
def test_oracle_mma_matches_reference_mma(oracle, mma_ref, m):
    rng = np.random.default_rng(21)
    n = 500
    mo, mr = oracle.MMAOracle(n, m), mma_ref.MMARef(n, m)
    x = np.full(n, 0.3)
    xold_r = x.copy()
    for it in range(8):
        dfdx, gx, dgdx = problem(n, m, x, rng)
        lo, hi = mo.set_outer_movelimit(0.0, 1.0, 0.2, x)
        lo_r, hi_r = mr.set_outer_movelimit(0.0, 1.0, 0.2, x)
        assert np.array_equal(lo, lo_r) and np.array_equal(hi, hi_r)
        xo = mo.update(x, dfdx, gx, dgdx, lo, hi)
        xr = mr.update(x, dfdx, gx, dgdx, lo_r, hi_r)
        assert np.abs(xo - xr).max() <= 1e-10, (it, np.abs(xo - xr).max())
        ch = mr.design_change(xr, xold_r)
        assert abs(ch - np.abs(xr - x).max()) <= 1e-15
        x = xr.copy()
    mr.close()


@pytest.mark.gpu
@pytest.mark.parametrize("m", [1, 2])
def test_device_mma_matches_reference_mma(mma_ref, m):
    from topopt_in_petsc_b200 import Context, MMA
    rng = np.random.default_rng(5)
    ctx = Context(21, 11, 11)
    n = ctx.nel
    mg, mr = MMA(ctx, m), mma_ref.MMARef(n, m)
    x = np.full(n, 0.3)
    for it in range(8):
        dfdx, gx, dgdx = problem(n, m, x, rng)
        lo, hi = mr.set_outer_movelimit(0.0, 1.0, 0.2, x)
        xr = mr.update(x, dfdx, gx, dgdx, lo, hi)
        xg, lo_g, hi_g = x.copy(), np.empty(n), np.empty(n)
        mg.SetOuterMovelimit(0.0, 1.0, 0.2, xg, lo_g, hi_g)
        mg.Update(xg, dfdx, gx, [d.copy() for d in dgdx], lo_g, hi_g)
        assert np.abs(xg - xr).max() <= 1e-9, (it, np.abs(xg - xr).max())
        x = xr.copy()
    mg.close(); mr.close(); ctx.close()
