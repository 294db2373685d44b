import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
shape = (34, 21, 13)
ctx = Context(*shape)
phys = LinearElasticity(ctx, opts=solver_opts(spmv_kernel=k))
x = np.full(ctx.nel, 0.12)
phys.SetDensity(x, 1e-9, 1.0, 3.0)
ctx.sync(); print("density ok", flush=True)
u = np.random.default_rng(0).uniform(-1, 1, ctx.ndof)
y = phys.SpMV(u)
print("spmv ok kernel", k, float(np.abs(y).max()), flush=True)
