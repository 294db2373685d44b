"""CPU, world_size-2 (gloo): the host-side logic of the single-reduction (Chronopoulos-Gear) Jacobi-PCG the product runs on
several GPUs (csrc/cgsr.cuh): per iteration ONE operator application, ONE all-reduce of three sums (r.u, r.r, w.u), the operand
u = D^-1 r exchanged as one halo plane per neighbour.  Restated with the numpy oracle on two gloo ranks (partition from the
C-ABI) and compared with a direct solve and with the classical two-reduction PCG iteration count."""
import os
import sys

import numpy as np
import pytest

from test_slab_gloo import _free_port, _partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, shape, out_q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import topopt_oracle as O
    from topopt_in_petsc_b200 import _lib

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lib = _lib.load()
        nxn, nyn, nzn = shape
        g = O.Grid(nxn, nyn, nzn)
        KE = O.hex8_ke(g.dx, g.dy, g.dz, g.nu)
        x = np.random.default_rng(5).uniform(0.05, 1.0, g.nel)
        E = O.simp_modulus(x, 1e-9, 1.0, 3.0)
        N, RHS = O.cantilever_bc_load(g)
        K = O.assemble_K(g, KE, x, 1e-9, 1.0, 3.0, N)
        b = N * RHS
        kn0, kn1, ke0, ke1 = _partition(lib, nzn, world, rank)
        plane, eplane = 3 * nxn * nyn, g.nex * g.ney
        gn0, gn1 = (kn0 - 1 if kn0 > 0 else 0), (kn1 + 1 if kn1 < nzn else nzn)
        gl = O.Grid(nxn, nyn, gn1 - gn0, xc=(0.0, g.dx * g.nex, 0.0, g.dy * g.ney, 0.0, g.dz * (gn1 - gn0 - 1)))
        own = slice((kn0 - gn0) * plane, (kn1 - gn0) * plane)
        dloc = K.diagonal()[kn0 * plane:kn1 * plane]

        def apply(u_own):  # halo exchange of the operand + local matrix-free operator on the stored sub-grid
            u_loc = np.zeros((gn1 - gn0) * plane)
            u_loc[own] = u_own
            reqs, got = [], {}
            for nb, sl_send, key in ((rank + 1, slice(len(u_own) - plane, len(u_own)), "up"), (rank - 1, slice(0, plane), "dn")):
                if 0 <= nb < world:
                    reqs.append(dist.isend(torch.from_numpy(u_own[sl_send].copy()), nb))
                    got[key] = torch.empty(plane, dtype=torch.float64)
                    reqs.append(dist.irecv(got[key], nb))
            for r in reqs:
                r.wait()
            if "up" in got:
                u_loc[(kn1 - gn0) * plane:(kn1 + 1 - gn0) * plane] = got["up"].numpy()
            if "dn" in got:
                u_loc[:plane] = got["dn"].numpy()
            y = O.spmv_matrix_free(gl, KE, E[gn0 * eplane:(gn1 - 1) * eplane], N[gn0 * plane:gn1 * plane], u_loc)
            return y[own]

        def allsum(*vals):
            t = torch.tensor(list(vals), dtype=torch.float64)
            dist.all_reduce(t)
            return t.tolist()

        bo = b[kn0 * plane:kn1 * plane]
        bnorm = np.sqrt(allsum(float(bo @ bo))[0])
        # ---- single-reduction iteration (cgsr.cuh): U then S, one all-reduce of three sums -------------------------------
        # as the kernel runs it (cgsr.cu): the inverse diagonal is the fp32 copy, the residual is not stored (u = D^-1 r carries
        # it: u -= alpha D^-1 s, r = u / dinv inside the sums) and x is updated every other iteration (lazy update)
        dinv = (1.0 / dloc).astype(np.float32).astype(np.float64)
        xs = np.zeros_like(bo); u = dinv * bo; w = apply(u)
        p = np.zeros_like(bo); s = np.zeros_like(bo)
        rloc = u / dinv
        gam, rr, dlt = allsum(float(rloc @ u), float(rloc @ rloc), float(w @ u))
        gam_prev, alpha_prev, first, its_sr, nred_sr = 1.0, 1.0, True, 0, 1
        while np.sqrt(rr) > 1e-8 * bnorm and its_sr < 5000:
            for xm in (1, 2):  # iterations come in pairs; the even one leaves x alone
                if first:
                    beta, alpha = 0.0, gam / dlt
                else:
                    beta = gam / gam_prev
                    alpha = gam / (dlt - beta * gam / alpha_prev)
                p_old = p
                p = u + beta * p; s = w + beta * s
                if xm == 2:
                    xs += alpha_prev * p_old + alpha * p
                u = u - alpha * dinv * s
                w = apply(u)
                gam_prev, alpha_prev, first = gam, alpha, False
                rloc = u / dinv
                gam, rr, dlt = allsum(float(rloc @ u), float(rloc @ rloc), float(w @ u))
                its_sr += 1; nred_sr += 1
        # ---- classical PCG: two reductions per iteration ---------------------------------------------------------------------
        xc = np.zeros_like(bo); r = bo.copy(); z = r / dloc; p = z.copy()
        (rz,) = allsum(float(r @ z)); its_cl, nred_cl, rr2 = 0, 1, float("inf")
        while np.sqrt(rr2) > 1e-8 * bnorm and its_cl < 5000:
            q = apply(p)
            (pq,) = allsum(float(p @ q))
            a = rz / pq
            xc += a * p; r -= a * q; z = r / dloc
            rz_new, rr2 = allsum(float(r @ z), float(r @ r))
            p = z + (rz_new / rz) * p; rz = rz_new
            its_cl += 1; nred_cl += 2
        import scipy.sparse.linalg as spla
        U = spla.spsolve(K.tocsc(), b)[kn0 * plane:kn1 * plane]
        scale = float(np.abs(spla.spsolve(K.tocsc(), b)).max())
        out_q.put((rank, float(np.abs(xs - U).max()) / scale, float(np.abs(xc - U).max()) / scale, its_sr, its_cl, nred_sr, nred_cl))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(9, 5, 9)])
def test_two_rank_single_reduction_pcg(shape):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err_sr, err_cl, its_sr, its_cl, nred_sr, nred_cl in res:
        assert err_sr <= 1e-6 and err_cl <= 1e-6, (rank, err_sr, err_cl)
        assert abs(its_sr - its_cl) <= max(3, its_cl // 20), (its_sr, its_cl)  # same Krylov method, different recurrences
        assert nred_sr == its_sr + 1 and nred_cl == 2 * its_cl + 1              # one all-reduce per iteration instead of two
    assert res[0][3] == res[1][3]  # both ranks took the same decisions (identical global sums)
