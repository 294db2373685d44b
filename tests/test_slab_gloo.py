"""CPU, world_size-2 (gloo): the host-side z-slab decomposition of the path.

The product shards the state solve by element layers (tob_slab_partition, the rule of
LinearElasticity.cc:804-818), keeps one ghost node plane per neighbour plus a redundant ghost element layer,
and needs one halo plane per operator application and one scalar all-reduce per dot product.  This test
restates that scheme with the numpy oracle on two gloo ranks -- partition from the C-ABI, halo planes by
send/recv, dots by all_reduce -- and checks it reproduces the undecomposed operator.  (The CUDA/NVLink
version of the same exchange is covered by tests/test_multigpu_gpu.py on 2 GPUs.)
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partition(lib, nzn, nranks, rank):
    v = [C.c_int() for _ in range(4)]
    rc = lib.tob_slab_partition(nzn, nranks, rank, *[C.byref(x) for x in v])
    assert rc == 0
    return tuple(x.value for x in v)


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
        rng = np.random.default_rng(7)  # same global problem on every rank
        x = rng.uniform(0.05, 1.0, g.nel)
        E = O.simp_modulus(x, 1e-9, 1.0, 3.0)
        N, _ = O.cantilever_bc_load(g)
        u = rng.standard_normal(g.ndof)
        y_ref = O.spmv_matrix_free(g, KE, E, N, u)

        kn0, kn1, ke0, ke1 = _partition(lib, nzn, world, rank)
        plane = 3 * nxn * nyn
        eplane = g.nex * g.ney
        # stored ranges: one ghost node plane per neighbour, one ghost element layer above
        gn0 = kn0 - 1 if kn0 > 0 else 0
        gn1 = kn1 + 1 if kn1 < nzn else nzn
        u_loc = np.zeros((gn1 - gn0) * plane)
        u_loc[(kn0 - gn0) * plane:(kn1 - gn0) * plane] = u[kn0 * plane:kn1 * plane]  # owned planes only

        # halo: first owned plane down, last owned plane up
        reqs = []
        if rank + 1 < world:
            up = torch.from_numpy(u[(kn1 - 1) * plane:kn1 * plane].copy())
            reqs.append(dist.isend(up, rank + 1))
            got_up = torch.empty(plane, dtype=torch.float64)
            reqs.append(dist.irecv(got_up, rank + 1))
        if rank > 0:
            dn = torch.from_numpy(u[kn0 * plane:(kn0 + 1) * plane].copy())
            reqs.append(dist.isend(dn, rank - 1))
            got_dn = torch.empty(plane, dtype=torch.float64)
            reqs.append(dist.irecv(got_dn, rank - 1))
        for r in reqs:
            r.wait()
        if rank + 1 < world:
            u_loc[(kn1 - gn0) * plane:(kn1 + 1 - gn0) * plane] = got_up.numpy()
        if rank > 0:
            u_loc[:plane] = got_dn.numpy()

        # local operator on the stored sub-grid [gn0, gn1) with element layers [gn0, gn1-1)
        gl = O.Grid(nxn, nyn, gn1 - gn0,
                    xc=(0.0, g.dx * g.nex, 0.0, g.dy * g.ney, 0.0, g.dz * (gn1 - gn0 - 1)))
        y_loc = O.spmv_matrix_free(gl, KE, E[gn0 * eplane:(gn1 - 1) * eplane],
                                   N[gn0 * plane:gn1 * plane], u_loc)
        y_own = y_loc[(kn0 - gn0) * plane:(kn1 - gn0) * plane]
        err = float(np.max(np.abs(y_own - y_ref[kn0 * plane:kn1 * plane])))

        # distributed dot product over owned ranges == global dot
        d = torch.tensor([float(u[kn0 * plane:kn1 * plane] @ y_own)], dtype=torch.float64)
        dist.all_reduce(d)
        derr = abs(d.item() - float(u @ y_ref))

        # ownership tiles the grid
        own = torch.tensor([kn1 - kn0, ke1 - ke0], dtype=torch.int64)
        dist.all_reduce(own)
        out_q.put((rank, err, derr, own.tolist(), (kn0, kn1, ke0, ke1), float(np.max(np.abs(y_ref)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(7, 5, 10), (5, 4, 4)])
def test_two_rank_slab_operator_matches_global(shape):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    nzn = shape[2]
    for rank, err, derr, own, slab, scale in res:
        assert err <= 1e-13 * scale, (rank, err)
        assert derr <= 1e-11 * scale * scale * 100, (rank, derr)
        assert own == [nzn, nzn - 1]
    slabs = dict((r[0], r[4]) for r in res)
    assert slabs[0][0] == 0 and slabs[0][1] == slabs[1][0] and slabs[1][1] == nzn  # node planes contiguous
    assert slabs[0][2] == 0 and slabs[0][3] == slabs[1][2] and slabs[1][3] == nzn - 1  # element layers contiguous


def test_partition_rule_many_ranks():
    from topopt_in_petsc_b200 import _lib
    lib = _lib.load()
    for nzn in (2, 3, 9, 97, 193):
        for nr in (1, 2, 3, 4, 8):
            if nzn - 1 < nr:
                v = [C.c_int() for _ in range(4)]
                assert lib.tob_slab_partition(nzn, nr, 0, *[C.byref(x) for x in v]) != 0
                continue
            prev_n, prev_e = 0, 0
            for r in range(nr):
                kn0, kn1, ke0, ke1 = _partition(lib, nzn, nr, r)
                assert kn0 == prev_n and ke0 == prev_e and kn1 > kn0 and ke1 > ke0
                assert ke1 == kn1 - 1  # an element follows its upper node plane
                prev_n, prev_e = kn1, ke1
            assert prev_n == nzn and prev_e == nzn - 1
