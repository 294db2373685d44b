"""GMG parameter sweep on the cfg2 cantilever: levels x coarse iterations x Chebyshev degree."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
size = sys.argv[1] if len(sys.argv) > 1 else "192x96x96"
ex, ey, ez = [int(v) for v in size.split("x")]
ctx = Context(ex + 1, ey + 1, ez + 1)
rng = np.random.default_rng(0)
xs = {"uniform": np.full(ctx.nel, 0.12)}
i, j, k = np.meshgrid(np.arange(ex), np.arange(ey), np.arange(ez), indexing="ij")
beam = ((np.abs(j - ey / 2) < ey / 8) & (np.abs(k - ez / 2) < ez / 8)).astype(float).transpose(2, 1, 0).ravel()
xs["beam_0/1"] = 0.001 + 0.998 * beam
for nl, coarse in ((4, 60), (5, 30), (5, 15), (6, 20), (6, 10)):
    for cheb in (3, 4):
        for name, x in xs.items():
            mg = LinearElasticity(ctx, opts=solver_opts(pc=1, nlvls=nl, rtol=1e-5, maxits=300, cheb_degree=cheb, coarse_its=coarse))
            s = mg.SolveState(x, 1e-9, 1.0, 3.0)   # includes the eager + capture V-cycles
            U = mg.GetStateField(); mg.SetStateField(np.zeros_like(U))
            s2 = mg.SolveState(x, 1e-9, 1.0, 3.0)  # graph replay only
            print(json.dumps({"size": size, "nlvls": nl, "coarse_its": coarse, "cheb": cheb, "design": name, "its": s2.iterations,
                              "ms": round(s2.seconds * 1e3, 2), "ms_per_it": round(s2.seconds * 1e3 / max(s2.iterations, 1), 3)}), flush=True)
            mg.close()
