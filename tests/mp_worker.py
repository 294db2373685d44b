"""Worker for tests/test_multigpu_gpu.py (launched by torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from topopt_in_petsc_b200 import Context, Filter, LinearElasticity, MMA, nccl_unique_id, solver_opts  # noqa: E402
import topopt_oracle as o  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)
shape = (21, 13, 19)
g = o.Grid(*shape)
ctx = Context(*shape, rank=rank, nranks=world, device=local, nccl_id=bytes(idt.cpu().tolist()))
nex, ney = g.nex, g.ney
e0, e1 = ctx.ke0 * nex * ney, ctx.ke1 * nex * ney
d0, d1 = 3 * ctx.kn0 * g.nxn * g.nyn, 3 * ctx.kn1 * g.nxn * g.nyn
assert ctx.nel_owned == e1 - e0 and ctx.ndof_owned == d1 - d0 and ctx.nel == g.nel
rng = np.random.default_rng(3)
x = rng.uniform(0.05, 1.0, g.nel)
u = rng.uniform(-1, 1, g.ndof)
KE = o.hex8_ke(g.dx, g.dy, g.dz, 0.3)
N, RHS = o.cantilever_bc_load(g)
E = o.simp_modulus(x, 1e-9, 1.0, 3.0)
# SpMV
phys = LinearElasticity(ctx, opts=solver_opts(rtol=1e-11, maxits=50000))
phys.SetDensity(x[e0:e1].copy(), 1e-9, 1.0, 3.0)
y = phys.SpMV(u[d0:d1].copy())
yref = o.spmv_matrix_free(g, KE, E, N, u)
assert np.abs(y - yref[d0:d1]).max() <= 1e-12 * np.abs(yref).max(), "spmv"
# solve + objective + sensitivities
K = o.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
import scipy.sparse.linalg as spla
U = spla.spsolve(K.tocsc(), N * RHS)
fx0, gx0, df0, dg0 = o.objective_sens(g, KE, U, x, 1e-9, 1.0, 3.0, 0.12)
df, dg = np.empty(ctx.nel_owned), np.empty(ctx.nel_owned)
fx, gx = phys.ComputeObjectiveConstraintsSensitivities(df, dg, x[e0:e1].copy(), 1e-9, 1.0, 3.0, 0.12)
assert abs(fx - fx0) <= 1e-7 * abs(fx0) and abs(gx - gx0) <= 1e-12, ("obj", fx, fx0)
assert np.abs(df - df0[e0:e1]).max() <= 1e-6 * np.abs(df0).max(), "sens"
assert np.abs(phys.GetStateField() - U[d0:d1]).max() <= 1e-6 * np.abs(U).max(), "state"
# single-reduction Jacobi-PCG (cgsr.cuh; the default on several ranks for rtol >= 1e-8): halo planes stored by the
# neighbours' update kernels, sums published by the last CTA of the SpMV
psr = LinearElasticity(ctx, opts=solver_opts(rtol=1e-8, maxits=50000, check_every=10))
ssr = psr.SolveState(x[e0:e1].copy(), 1e-9, 1.0, 3.0)
assert ssr.rel_residual <= 1e-8 and ssr.iterations > 10, ("sr-pcg", ssr.rel_residual, ssr.iterations)
assert np.abs(psr.GetStateField() - U[d0:d1]).max() <= 1e-5 * np.abs(U).max(), "sr-pcg state"
ssr2 = psr.SolveState((0.8 * x[e0:e1] + 0.1).copy(), 1e-9, 1.0, 3.0)  # second solve: replays the graph, epochs continue
assert ssr2.rel_residual <= 1e-8
if rank == 0:
    print("sr-pcg iterations", ssr.iterations, ssr2.iterations, flush=True)
psr.close()
# filter (2-cell radius so ghost layers matter)
fo = o.FilterOracle(g, 1, 0.25)
flt = Filter(ctx, 1, 0.25)
xt, xp = np.empty(ctx.nel_owned), np.empty(ctx.nel_owned)
flt.FilterProject(x[e0:e1].copy(), xt, xp, True, 4.0, 0.3)
xt0, xp0 = fo.filter_project(x, True, 4.0, 0.3)
assert np.abs(xt - xt0[e0:e1]).max() <= 1e-13 and np.abs(xp - xp0[e0:e1]).max() <= 1e-13, "filter"
gf, gg = df0[e0:e1].copy(), dg0[e0:e1].copy()
flt.Gradients(x[e0:e1].copy(), xt, gf, 1, [gg], True, 4.0, 0.3)
gf0, gg0 = fo.gradients(x, xt0, df0, [dg0], True, 4.0, 0.3)
assert np.abs(gf - gf0[e0:e1]).max() <= 1e-12 * np.abs(gf0).max(), "filter grad"
assert abs(flt.GetMND(xp) - o.mnd(xp0)) <= 1e-13, "mnd"
# PDE filter across the slab boundary
po = o.PDEFilterOracle(g, 0.3)
fp = Filter(ctx, 2, 0.3)
out = fp.PDEApply(x[e0:e1].copy())
assert np.abs(out - po.apply(x)[e0:e1]).max() <= 1e-7, "pde filter"
fp.close()
# multigrid-preconditioned solve on 2 slabs (16 element layers per rank, 3 levels, aligned slab boundaries)
# multigrid: (17,9,33) has slab boundaries aligned for all 3 levels, (17,9,29) only for 2 (14 layers per rank): the
# levels below the last aligned / large-enough one are replicated on every rank (gmg.cu), so both must work and
# converge like a single-rank V-cycle
for shape2 in ((17, 9, 33), (17, 9, 29)):
    g2 = o.Grid(*shape2)
    idt2 = torch.zeros(128, dtype=torch.uint8, device="cuda")  # every communicator needs its own ncclUniqueId
    if rank == 0:
        idt2 = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt2, 0)
    ctx2 = Context(*shape2, rank=rank, nranks=world, device=local, nccl_id=bytes(idt2.cpu().tolist()))
    x2 = np.random.default_rng(8).uniform(0.05, 1.0, g2.nel)
    KE2 = o.hex8_ke(g2.dx, g2.dy, g2.dz, 0.3)
    N2, RHS2 = o.cantilever_bc_load(g2)
    U2 = spla.spsolve(o.assemble_K(g2, KE2, x2, 1e-9, 1.0, 3.0, N2).tocsc(), N2 * RHS2)
    a0, a1 = ctx2.ke0 * g2.nex * g2.ney, ctx2.ke1 * g2.nex * g2.ney
    b0, b1 = 3 * ctx2.kn0 * g2.nxn * g2.nyn, 3 * ctx2.kn1 * g2.nxn * g2.nyn
    pmg = LinearElasticity(ctx2, opts=solver_opts(pc=1, nlvls=3, rtol=1e-9, maxits=300))
    st = pmg.SolveState(x2[a0:a1].copy(), 1e-9, 1.0, 3.0)
    assert st.rel_residual <= 1e-9 and st.iterations < 100, ("gmg its", shape2, st.iterations)
    assert np.abs(pmg.GetStateField() - U2[b0:b1]).max() <= 1e-6 * np.abs(U2).max(), ("gmg state", shape2)
    st2 = pmg.SolveState((0.9 * x2[a0:a1] + 0.05).copy(), 1e-9, 1.0, 3.0)  # second solve replays the V-cycle graph
    assert st2.rel_residual <= 1e-9
    if rank == 0:
        print("gmg", shape2, "iterations", st.iterations, st2.iterations, flush=True)
    pmg.close()
    # the same solve with the finest-level smoother in single precision on the tensor cores (TF32; the interleaved fp32 iterate is
    # halo-exchanged between the two slabs): same answer, about the same number of outer iterations
    os.environ["TOB_GMG_TF32"] = "1"
    ptf = LinearElasticity(ctx2, opts=solver_opts(pc=1, nlvls=3, rtol=1e-9, maxits=300))
    stf = ptf.SolveState(x2[a0:a1].copy(), 1e-9, 1.0, 3.0)
    os.environ["TOB_GMG_TF32"] = "0"
    assert stf.rel_residual <= 1e-9 and stf.iterations <= st.iterations + 8, ("tf32 smoother", shape2, stf.iterations, st.iterations)
    assert np.abs(ptf.GetStateField() - U2[b0:b1]).max() <= 1e-6 * np.abs(U2).max(), ("tf32 smoother state", shape2)
    if rank == 0:
        print("gmg + tf32 smoother", shape2, "iterations", stf.iterations, flush=True)
    ptf.close(); ctx2.close()
# MMA
mo, mg = o.MMAOracle(g.nel, 1), MMA(ctx, 1)
xm = np.full(g.nel, 0.3)
for it in range(4):
    c0 = np.random.default_rng(10 + it).uniform(0.5, 2.0, g.nel)
    dfx = -c0 / (xm + 0.1) ** 2
    gxv = np.array([xm.mean() - 0.3])
    dgx = np.full(g.nel, 1.0 / g.nel)
    lo, hi = mo.set_outer_movelimit(0.0, 1.0, 0.2, xm)
    xn = mo.update(xm, dfx, gxv, [dgx], lo, hi)
    xl, lo_l, hi_l = xm[e0:e1].copy(), np.empty(e1 - e0), np.empty(e1 - e0)
    mg.SetOuterMovelimit(0.0, 1.0, 0.2, xl, lo_l, hi_l)
    mg.Update(xl, dfx[e0:e1].copy(), gxv, [dgx[e0:e1].copy()], lo_l, hi_l)
    assert np.abs(xl - xn[e0:e1]).max() <= 1e-9, ("mma", it)
    xm = xn
for obj in (mg, flt, phys):
    obj.close()
ctx.close()
dist.barrier()
if rank == 0:
    print("MP_PARITY_OK")
dist.destroy_process_group()
