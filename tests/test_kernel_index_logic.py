"""Host-side restatements of two pieces of index logic the CUDA kernels rely on -- no GPU needed.

1. csrc/cgsr.cu, k_cg_update: the update kernel of the multi-GPU PCG processes the two boundary planes of the owned slab first
   (phase A, by the first `nbnd` blocks, which then raise the neighbours' flags) and the interior afterwards (phase B, grid-stride,
   block order rotated).  Every index must be visited exactly once, and no block outside the first `nbnd` may touch a boundary
   plane -- otherwise a neighbour could read a ghost plane that is still being written.
2. csrc/hex8_tc32.cu, k_hex8_tc32c<2>: the node sums of the tensor-core operator are formed with warp shuffles (x neighbour =
   lane - 1, y neighbour = lane - 8) and a small shared-memory hand-over for the row that crosses a warp boundary.  Emulated
   here with numpy against the direct sum over the four elements around a node."""
import numpy as np
import pytest


def _visits(n2, plane2, G, B):
    nA = min(2 * plane2, n2)
    nbnd = min(G, -(-nA // B))
    seen = np.zeros(n2, dtype=np.int32)
    owner = np.full(n2, -1, dtype=np.int64)
    for b in range(G):
        if b < nbnd:
            j = np.arange(b * B, nA, nbnd * B)[:, None] + np.arange(B)[None, :]
            j = j[j < nA]
            i = np.where(j < plane2, j, n2 - nA + j)
            np.add.at(seen, i, 1)
            owner[i] = b
        nI = n2 - nA
        bo = (b + G - nbnd) % G
        k = np.arange(bo * B, max(nI, 0), G * B)[:, None] + np.arange(B)[None, :]
        k = k[k < nI]
        np.add.at(seen, plane2 + k, 1)
        owner[plane2 + k] = b
    return seen, owner, nbnd


@pytest.mark.parametrize("n2,plane2,G,B", [
    (193 * 97 * 3 * 49 // 2, 193 * 97 * 3 // 2, 592, 256),   # cfg2 on 2 ranks
    (193 * 97 * 3 * 12 // 2, 193 * 97 * 3 // 2, 592, 256),   # cfg2 on 8 ranks
    (68 * 5, 68, 3, 64),        # tiny grid, fewer blocks than boundary chunks
    (68 * 2, 68, 1, 64),        # two owned planes, one block
    (68, 68, 4, 32),            # a single owned plane: both roles on the same indices
    (1000, 7, 592, 256),        # boundary smaller than a block
])
def test_update_kernel_visits_every_index_once_boundary_first(n2, plane2, G, B):
    seen, owner, nbnd = _visits(n2, plane2, G, B)
    assert (seen == 1).all()
    lo = owner[:plane2]
    hi = owner[max(n2 - plane2, 0):]
    assert (lo < nbnd).all() and (hi < nbnd).all()  # only phase-A blocks write what the neighbours read


def test_tensor_core_node_sums_by_shuffles():
    rng = np.random.default_rng(3)
    nthreads = 128
    Y = rng.standard_normal((nthreads, 8, 3))        # my element's 8 local-node outputs (already scaled by E)
    nloc = lambda a, b, g: (4 if g else 0) + ((2 if a else 3) if b else (1 if a else 0))
    t = np.arange(nthreads)
    ix, iy = t & 7, t >> 3
    out = (ix >= 1) & (iy >= 1)
    # direct: node thread t sums local node (a, b, g) of element t - (8 b + a)
    direct = np.zeros((nthreads, 2, 3))
    for tt in t[out]:
        for a in range(2):
            for b in range(2):
                e = tt - (8 * b + a)
                for g in range(2):
                    direct[tt, g] += Y[e, nloc(a, b, g)]
    # shuffles: per warp of 32 lanes
    shfl_up = lambda v, d: np.concatenate([v[:d], v[:-d]])  # lanes < d keep their own value
    got = np.zeros((nthreads, 2, 3))
    X = np.zeros((4, 6, 8))  # crossing rows: [warp][g * 3 + c][lane - 24]
    a1_all = np.zeros((nthreads, 6))
    for w in range(4):
        sl = slice(32 * w, 32 * w + 32)
        for g in range(2):
            for c in range(3):
                a0 = Y[sl, nloc(0, 0, g), c] + shfl_up(Y[sl, nloc(1, 0, g), c], 1)
                a1 = Y[sl, nloc(0, 1, g), c] + shfl_up(Y[sl, nloc(1, 1, g), c], 1)
                got[sl, g, c] = a0
                a1_all[sl, 3 * g + c] = a1
        if w < 3:
            X[w] = a1_all[sl][24:32].T
    for w in range(4):
        sl = slice(32 * w, 32 * w + 32)
        for q in range(6):
            bsum = shfl_up(a1_all[sl, q], 8)
            if w > 0:
                bsum[:8] = X[w - 1, q]
            else:
                bsum[:8] = 0.0
            got[sl, q // 3, q % 3] += bsum
    assert np.allclose(got[out], direct[out], rtol=0, atol=1e-12)
