"""CPU: the algebra the z-march SpMV kernel (csrc/hex8.cu) relies on, restated in numpy.

With the 8 element nodes indexed by their binary offsets q = a + 2b + 4g, W the 8x8 Walsh-Hadamard matrix over q and
T = W (x) I3, the box-element stiffness satisfies KE = T^T D T with D = T KE T^T / 64, and D is block diagonal:
(mode m, component c) couples with (m', c') only if m ^ (1 << c) == m' ^ (1 << c') =: rho, i.e. 8 symmetric 3x3 blocks,
of which rho = 1, 2, 4 contain the translation mode m = 0 and therefore a zero row/column.  The kernel applies T stage by
stage, the 8 blocks, and T^T; here the identity is checked on the oracle's KE (any dx, dy, dz, nu) -- the CUDA result
itself is compared with the assembled operator on the GPU (tests/test_hex8_gpu.py)."""
import numpy as np
import pytest

BIN2REF = [0, 1, 3, 2, 4, 5, 7, 6]  # binary node index q -> reference local node (LinearElasticity.cc:118-120 order)


def binary_order_ke(KE):
    idx = [3 * BIN2REF[q] + c for q in range(8) for c in range(3)]
    return KE[np.ix_(idx, idx)]


def hadamard8():
    return np.array([[(-1) ** bin(m & q).count("1") for q in range(8)] for m in range(8)], dtype=float)


@pytest.mark.parametrize("d", [(0.05, 0.05, 0.05), (2.0 / 12, 1.0 / 8, 1.0 / 6), (1.0, 0.3, 2.5)])
@pytest.mark.parametrize("nu", [0.3, 0.12])
def test_block_diagonal_form_and_reconstruction(oracle, d, nu):
    KE = binary_order_ke(oracle.hex8_ke(*d, nu))
    T = np.kron(hadamard8(), np.eye(3))
    D = T @ KE @ T.T / 64.0
    scale = np.abs(D).max()
    # coupling pattern: only entries with equal rho survive
    rho = np.array([m ^ (1 << c) for m in range(8) for c in range(3)])
    off = D[rho[:, None] != rho[None, :]]
    assert np.abs(off).max() <= 1e-13 * scale
    blocks = {}
    for r in range(8):
        sel = np.where(rho == r)[0]
        assert len(sel) == 3
        B = D[np.ix_(sel, sel)]
        assert np.abs(B - B.T).max() <= 1e-14 * scale
        blocks[r] = B
        comps = [int(i) % 3 for i in sel]
        modes = [int(i) // 3 for i in sel]
        for k, (m, c) in enumerate(zip(modes, comps)):
            assert m == r ^ (1 << c)
            if m == 0:  # rigid translation of component c: zero row and column (rho = 1 << c)
                assert np.abs(B[k]).max() <= 1e-13 * scale and np.abs(B[:, k]).max() <= 1e-13 * scale
    assert sum(1 for r in (1, 2, 4) if np.abs(blocks[r]).max() > 0) == 3
    # reconstruction: KE = T^T D T, applied the way the kernel does (transform, 8 blocks, transform back)
    u = np.random.default_rng(0).standard_normal(24)
    uh = T @ u
    yh = np.zeros(24)
    for r in range(8):
        sel = np.where(rho == r)[0]
        yh[sel] = blocks[r] @ uh[sel]
    assert np.abs(T.T @ yh - KE @ u).max() <= 1e-13 * np.abs(KE @ u).max()
    # operation count quoted in DESIGN.md: 5 full blocks (3 mul + 6 fma) + 3 blocks without the translation row/column
    assert 5 * 9 + 3 * 4 == 57
