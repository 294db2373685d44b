"""One-GPU stand-in for a rank's share of the cfg2 slab at N = 1, 2, 4, 8: the single-reduction PCG (TOB_PCG_SR=1) on
192 x 96 x (96 / N) elements with the device time stamps (TOB_CG_TRACE=1) -- kernel bodies without any cross-GPU wait.
Usage: python scripts/sr_probe.py [nez ...]"""
import os, sys
os.environ["TOB_PCG_SR"] = "1"
os.environ["TOB_CG_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from topopt_in_petsc_b200 import Context

for nez in [int(a) for a in sys.argv[1:]] or [96, 48, 24, 12]:
    c = dict(bench.BASE, **bench.WORKLOADS["cfg2"])
    c["nez"] = nez
    c["xc"] = (0, 2, 0, 1, 0, nez / 96.0)
    ctx = Context(c["nex"] + 1, c["ney"] + 1, c["nez"] + 1, xc=c["xc"])
    loop = bench.DesignLoop(ctx, c, resident=True, profile=False, pc=0, nlvls=1)
    print(f"--- 192x96x{nez}", file=sys.stderr, flush=True)
    for _ in range(2):
        loop.step()
    ctx.sync()
    print(f"    its {loop.pcg_its}", file=sys.stderr, flush=True)
    loop.close(); ctx.close()
