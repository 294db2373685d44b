"""GMG-PCG iteration growth over the first design iterations of a workload, for a few hierarchy settings.
Usage: python scripts/gmg_growth_probe.py [cfg2|cfg3] [steps]   (env TOB_GMG_GALERKIN_FROM is set per run)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from topopt_in_petsc_b200 import Context

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = dict(bench.BASE, **bench.WORKLOADS[wl])
runs = [dict(gf="2", cheb=3), dict(gf="1", cheb=3), dict(gf="0", cheb=3), dict(gf="2", cheb=5), dict(gf="1", cheb=5)]
if len(sys.argv) > 3:
    runs = [dict(gf=a.split(":")[0], cheb=int(a.split(":")[1])) for a in sys.argv[3:]]
for r in runs:
    os.environ["TOB_GMG_GALERKIN_FROM"] = r["gf"]
    ctx = Context(c["nex"] + 1, c["ney"] + 1, c["nez"] + 1, xc=c["xc"])
    loop = bench.DesignLoop(ctx, c, resident=True, profile=False, pc=1, nlvls=bench.gmg_levels(c))
    loop.phys.SetOpts(cheb_degree=r["cheb"])
    t0 = time.time()
    for _ in range(steps):
        loop.step()
    ctx.sync()
    dt = time.time() - t0
    print(json.dumps({"workload": wl, "galerkin_from": r["gf"], "cheb": r["cheb"], "its": loop.pcg_its, "s_per_step": dt / steps,
                      "fx": loop.hist[-1]["fx"]}), flush=True)
    loop.close(); ctx.close()
