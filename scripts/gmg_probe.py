"""GMG vs Jacobi PCG: iterations and solve time on the cantilever for a few design states."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
size = sys.argv[1] if len(sys.argv) > 1 else "96x48x48"
ex, ey, ez = [int(v) for v in size.split("x")]
ctx = Context(ex + 1, ey + 1, ez + 1)
rng = np.random.default_rng(0)
xs = {"uniform": np.full(ctx.nel, 0.12), "random": rng.uniform(0.01, 1, ctx.nel)}
i, j, k = np.meshgrid(np.arange(ex), np.arange(ey), np.arange(ez), indexing="ij")
beam = ((np.abs(j - ey / 2) < ey / 8) & (np.abs(k - ez / 2) < ez / 8)).astype(float).transpose(2, 1, 0).ravel()
xs["beam_0/1"] = 0.001 + 0.998 * beam
for cheb in (2, 3, 4):
    for ratio_name, coarse in (("c30", 30), ("c60", 60)):
        for name, x in xs.items():
            mg = LinearElasticity(ctx, opts=solver_opts(pc=1, nlvls=4, rtol=1e-5, maxits=300, cheb_degree=cheb, coarse_its=coarse))
            s = mg.SolveState(x, 1e-9, 1.0, 3.0)
            print(json.dumps({"size": size, "pc": "gmg", "cheb": cheb, "coarse_its": coarse, "design": name, "its": s.iterations,
                              "relres": s.rel_residual, "ms": s.seconds * 1e3}), flush=True)
            mg.close()
for name, x in xs.items():
    jac = LinearElasticity(ctx, opts=solver_opts(pc=0, rtol=1e-5, maxits=100000, check_every=50))
    s = jac.SolveState(x, 1e-9, 1.0, 3.0)
    print(json.dumps({"size": size, "pc": "jacobi", "design": name, "its": s.iterations, "relres": s.rel_residual, "ms": s.seconds * 1e3}), flush=True)
    jac.close()
