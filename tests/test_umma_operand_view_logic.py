"""Host-side restatement of the operand view the tensor-core kernels rely on (csrc/hex8_tc32.cu, generations 2 and 3): with the
fp32 vector stored node-interleaved ("V4": one 16-byte chunk (ux, uy, uz, 0) per node) a plane tile in shared memory IS a
canonical K-major UMMA operand -- row r of the M = 128 tile and K-chunk a' sit at  start + (r // 8) * SBO + (r % 8) * 16 + a' * LBO
with LBO = 16 B (the two chunks of a K = 8 step overlap in memory) and SBO = one tile row.  This test emulates that address
arithmetic in numpy, multiplies with the operand blocks the host builds from KE (tc3_build_B / tc2_build_B restated), and checks
the result against KE u_e of the oracle -- no GPU needed."""
import numpy as np


def ref(a, b, g):  # reference local node of corner (a, b, g), LinearElasticity.cc:118-120
    return (4 if g else 0) + ((2 if a else 3) if b else (1 if a else 0))


def umma_rows(smem_bytes, start, lbo, sbo, m=128):
    """the [m][8] fp32 operand a K-major no-swizzle descriptor (start, LBO, SBO) addresses: chunk c in {0,1}, 4 floats each"""
    f = smem_bytes.view(np.float32)
    out = np.empty((m, 8), dtype=np.float32)
    for r in range(m):
        for c in range(2):
            o = (start + (r // 8) * sbo + (r % 8) * 16 + c * lbo) // 4
            out[r, 4 * c:4 * c + 4] = f[o:o + 4]
    return out


def test_generation3_element_rows(oracle):
    rng = np.random.default_rng(0)
    g = oracle.Grid(9, 17, 3)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3).astype(np.float32)
    PX, PY = 9, 17  # plane tile of 8 x 16 elements
    planes = [rng.uniform(-1, 1, (PY, PX, 4)).astype(np.float32) for _ in range(2)]
    for pl in planes:
        pl[..., 3] = 0.0
    smem = np.concatenate([pl.reshape(-1) for pl in planes]).view(np.uint8)
    slot = PX * PY * 16
    D = np.zeros((128, 32), dtype=np.float64)
    for st in range(4):  # K-step (b', g'); chunks a' = 0, 1
        bp, gp = st & 1, st >> 1
        A = umma_rows(smem, gp * slot + bp * PX * 16, 16, PX * 16)
        B = np.zeros((32, 8), dtype=np.float32)  # tc3_build_B: B[4 node + c][a' * 4 + c'] = KE[3 node + c][3 ref(a', b', g') + c']
        for ap in range(2):
            for nd in range(8):
                for c in range(3):
                    B[4 * nd + c, 4 * ap:4 * ap + 3] = KE[3 * nd + c, 3 * ref(ap, bp, gp):3 * ref(ap, bp, gp) + 3]
        D += A.astype(np.float64) @ B.astype(np.float64).T
    for e in (0, 5, 8, 77, 127):
        ex, ey = e % 8, e // 8
        ue = np.zeros(24)
        for gg in range(2):
            for b in range(2):
                for a in range(2):
                    ue[3 * ref(a, b, gg):3 * ref(a, b, gg) + 3] = planes[gg][ey + b, ex + a, :3]
        got = D[e].reshape(8, 4)
        assert np.abs(got[:, :3].reshape(-1) - KE.astype(np.float64) @ ue).max() <= 1e-5 and np.abs(got[:, 3]).max() == 0.0


def test_generation2_node_rows_with_shifts(oracle):
    rng = np.random.default_rng(1)
    g = oracle.Grid(9, 17, 3)
    KE = oracle.hex8_ke(g.dx, g.dy, g.dz, 0.3).astype(np.float64)
    PX, PY = 10, 18  # 8 x 16 nodes + halo
    planes = [rng.uniform(-1, 1, (PY, PX, 4)).astype(np.float32) for _ in range(3)]  # z - 1, z, z + 1
    for pl in planes:
        pl[..., 3] = 0.0
    smem = np.concatenate([pl.reshape(-1) for pl in planes]).view(np.uint8)
    slot = PX * PY * 16
    E = rng.uniform(0.1, 1.0, (2, PY, PX))  # modulus of element (layer, ey, ex), lower-corner indexing in tile coordinates
    for node in (0, 7, 64, 127):
        ix, iy = node % 8, node // 8
        y = np.zeros(3)
        yref = np.zeros(3)
        for s in range(8):
            a, b, gsh = s & 1, (s >> 1) & 1, s >> 2
            Ds = np.zeros(3)
            for st in range(4):
                bp, gp = st & 1, st >> 1
                pz = 1 - gsh + gp  # plane index in the 3-plane ring (z = 1)
                A = umma_rows(smem, pz * slot + ((1 - b + bp) * PX + (1 - a)) * 16, 16, PX * 16)
                for ap in range(2):
                    for n in range(3):
                        Ds[n] += A[node, 4 * ap:4 * ap + 3].astype(np.float64) @ KE[3 * ref(a, b, gsh) + n, 3 * ref(ap, bp, gp):3 * ref(ap, bp, gp) + 3]
            # element I - s: lower corner at tile node (ix + 1 - a, iy + 1 - b), layer z - g
            Ee = E[1 - gsh, iy + 1 - b, ix + 1 - a]
            y += Ee * Ds
            ue = np.zeros(24)
            for gg in range(2):
                for bb in range(2):
                    for aa in range(2):
                        ue[3 * ref(aa, bb, gg):3 * ref(aa, bb, gg) + 3] = planes[1 - gsh + gg][iy + 1 - b + bb, ix + 1 - a + aa, :3]
            yref += Ee * (KE @ ue)[3 * ref(a, b, gsh):3 * ref(a, b, gsh) + 3]
        assert np.abs(y - yref).max() <= 1e-5
