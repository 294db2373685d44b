"""CPU: the index logic of the TMA-tiled cone-filter kernel (csrc/filter.cu, k_cone_tiled) restated in numpy --
zero-filled box with an even-padded x rim, plane-wise scatter into the output layers, weights by |offset| -- against
the oracle's explicit filter matrix H (Filter.cc:404-448).  The CUDA kernel itself is checked on the GPU by
tests/test_filter_mma_gpu.py; this keeps its tiling arithmetic honest without one."""
import numpy as np
import pytest

CT_X, CT_Y, CT_Z = 32, 8, 8


def tiled_cone(field, w, s, l0, l1):
    """field[l, j, i] = filter-layout input (all stored layers); returns sum_w for owned layers [l0, l1)."""
    nl, ney, nex = field.shape
    sp = (s + 1) & ~1
    bw, bh, bd = CT_X + 2 * sp, CT_Y + 2 * s, CT_Z + 2 * s
    out = np.zeros((l1 - l0, ney, nex))
    for bz in range(-(-(l1 - l0) // CT_Z)):
        for by in range(-(-ney // CT_Y)):
            for bx in range(-(-nex // CT_X)):
                i0, j0, lz = bx * CT_X, by * CT_Y, l0 + bz * CT_Z
                box = np.zeros((bd, bh, bw))  # TMA box load with out-of-bounds zero fill
                for zp in range(bd):
                    for yy in range(bh):
                        l, j = lz - s + zp, j0 - s + yy
                        if not (0 <= l < nl and 0 <= j < ney):
                            continue
                        a, b = i0 - sp, i0 - sp + bw
                        lo, hi = max(a, 0), min(b, nex)
                        if lo < hi:
                            box[zp, yy, lo - a:hi - a] = field[l, j, lo:hi]
                acc = np.zeros((CT_Z, CT_Y, CT_X))
                for zp in range(bd):
                    for dj in range(-s, s + 1):
                        for di in range(-s, s + 1):
                            v = box[zp, s + dj:s + dj + CT_Y, sp + di:sp + di + CT_X]
                            for dk in range(-s, s + 1):
                                o = zp - s - dk
                                if 0 <= o < CT_Z:
                                    acc[o] += w[abs(dk), abs(dj), abs(di)] * v
                for o in range(CT_Z):
                    l = lz + o
                    if l >= l1:
                        break
                    jj, ii = min(CT_Y, ney - j0), min(CT_X, nex - i0)
                    out[l - l0, j0:j0 + jj, i0:i0 + ii] = acc[o, :jj, :ii]
    return out


@pytest.mark.parametrize("shape,rmin", [((41, 13, 11), 0.12), ((35, 10, 19), 0.08)])
def test_tiled_cone_matches_filter_matrix(oracle, shape, rmin):
    g = oracle.Grid(*shape)
    s = oracle.filter_elemconn(g, rmin)
    assert s in (1, 2)
    H, Hs, _ = oracle.filter_matrix(g, rmin)
    x = np.random.default_rng(5).uniform(0, 1, g.nel)
    ref = H @ x
    w = np.zeros((s + 1, s + 1, s + 1))
    for a in range(s + 1):
        for b in range(s + 1):
            for c in range(s + 1):
                d = np.sqrt((c * g.dx) ** 2 + (b * g.dy) ** 2 + (a * g.dz) ** 2)
                w[a, b, c] = rmin - d if d < rmin else 0.0
    field = x.reshape(g.nez, g.ney, g.nex)
    got = tiled_cone(field, w, s, 0, g.nez).ravel()
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    # a z-slab with ghost layers: owned layers [3, 9) of the stored range [3 - s, 9 + s)
    lo, hi = 3, 9
    sub = field[lo - s:hi + s]
    got2 = tiled_cone(sub, w, s, s, s + hi - lo).ravel()
    assert np.abs(got2 - ref.reshape(g.nez, -1)[lo:hi].ravel()).max() <= 1e-13 * np.abs(ref).max()
