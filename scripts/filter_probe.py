"""GPU probe: cone-filter apply time (FilterProject = layout copy + cone stencil + copy) for the TMA-tiled and the
direct-gather kernel.  Usage: [TOB_CONE_TILED=0] python scripts/filter_probe.py [ex ey ez rcells]"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from topopt_in_petsc_b200 import Context, Filter
from topopt_in_petsc_b200.api import DevVec

ex, ey, ez, rc = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (192, 96, 96, 3))]
ctx = Context(ex + 1, ey + 1, ez + 1)
dx = 2.0 / ex
flt = Filter(ctx, 1, rc * dx)
n = ctx.nel
x = DevVec(ctx, n).upload(np.random.default_rng(3).uniform(0, 1, n))
xt, xp = DevVec(ctx, n), DevVec(ctx, n)
for _ in range(3):
    flt.FilterProject(x.ptr, xt.ptr, xp.ptr, False, 1.0, 0.5)
ctx.sync()
reps = 50
t0 = time.perf_counter()
for _ in range(reps):
    flt.FilterProject(x.ptr, xt.ptr, xp.ptr, False, 1.0, 0.5)
ctx.sync()
dt = (time.perf_counter() - t0) / reps
print(json.dumps({"grid": [ex, ey, ez], "s": flt.ElemConn(), "tiled": os.environ.get("TOB_CONE_TILED", "1"),
                  "filter_project_us": dt * 1e6, "checksum": float(xt.download().sum())}))
