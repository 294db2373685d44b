"""oracle/cpu_design_loop.py (what bench.py's reference arm times: whole design iterations with the reference's FGMRES +
Galerkin-PCMG solver configuration, C/OpenMP) reproduces the numpy oracle's design trajectory when both solve tightly."""
import numpy as np


def test_cpu_design_loop_matches_numpy_oracle(oracle):
    import cpu_design_loop as cd
    g = oracle.Grid(17, 9, 9)
    ref = oracle.run_design_loop(g, 3, rmin=0.3, solve_tol=1e-11)
    L = cd.CpuDesignLoop(16, 8, 8, rmin=0.3, nlvls=3, rtol=1e-11, nthreads=2)
    for k in range(3):
        r = L.step()
        h = ref["hist"][k]
        assert abs(r["fx"] - h["fx"]) <= 1e-9 * h["fx"]
        assert abs(r["ch"] - h["ch"]) <= 1e-8
        assert 0 < r["krylov_its"] <= 40  # multigrid-preconditioned: tens of iterations, not thousands
        s = r["seconds"]
        assert s["total"] >= s["assembly"] + s["galerkin_setup"] + s["solve"]
    assert np.abs(L.x - ref["x"]).max() <= 1e-8
    L.close()


def test_reference_tolerance_needs_few_iterations(oracle):
    import cpu_design_loop as cd
    L = cd.CpuDesignLoop(32, 16, 16, rmin=0.15, nlvls=4, rtol=1e-5, nthreads=2)
    its = [L.step()["krylov_its"] for _ in range(2)]
    assert all(0 < i <= 25 for i in its), its
    L.close()
