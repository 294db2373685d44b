// MPIIO.cc -- writes the reference's binary result container.
// Layout (MPIIO.cc:293-371 header, :380-508 mesh, :511-714 fields; reader bin2vtu_v3.py:23-144):
//   info string + 0x01 | u64 nDom | u64 nPointsT[nDom] nCellsT[] nPFields[] nCFields[] | u64 nodesPerElement
//   | point-field names + 0x01 | cell-field names + 0x01 | f32 xyz per point | u64 connectivity (8/cell)
//   | u64 offsets | u64 cell types (12 = VTK hexahedron) | per dump: u64 iteration, f32 point fields
//   (field-major), f32 cell fields (field-major).
// The reference writes this file collectively with MPI_File_write_at_all; here every rank pwrite()s its own z-slab
// at the same offsets (one box, one file system): with z-slabs the natural order is the global order, so a rank's
// points, cells and field values are contiguous runs of the global arrays and the offsets follow from the
// ownership rule alone (tob_slab_partition), without any communication.  Rank 0 writes the header.
#include "MPIIO.h"
#include <fcntl.h>
#include <unistd.h>
#include <vector>
#include "../../include/topopt_b200.h"

typedef unsigned long u64;

static bool put(int fd, const void* p, size_t n, u64 off) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        ssize_t w = pwrite(fd, c, n, (off_t)off);
        if (w <= 0) return false;
        c += w; n -= (size_t)w; off += (u64)w;
    }
    return true;
}

MPIIO::MPIIO(DM da_nodes, int nPfields, std::string pnames, int nCfields, std::string cnames)
    : nPoints(0), nCells(0), nPF(nPfields), nCF(nCfields), enabled(true), p0(0), c0(0), nPointsG(0), nCellsG(0), data0(0),
      ndumps(0) {
    PetscBool flg;
    char buf[PETSC_MAX_PATH_LEN];
    filename = "output_00000.dat";
    PetscOptionsGetString(NULL, NULL, "-workdir", buf, sizeof(buf), &flg);
    if (flg) filename = std::string(buf) + "/output.dat";  // MPIIO.cc:229-236
    PetscBool on = PETSC_TRUE;
    PetscOptionsGetBool(NULL, NULL, "-tob_output", &on, &flg);
    if (!on) enabled = false;
    if (!enabled) return;
    const int rank = TobRank();
    PetscInt n[3]; PetscScalar xc[6];
    DMGetGridTob(da_nodes, n, xc);
    int kn0, kn1, ke0, ke1;
    tob_ctx_slab(DMGetTobCtx(da_nodes), &kn0, &kn1, &ke0, &ke1);
    const u64 nplane = (u64)n[0] * n[1], eplane = (u64)(n[0] - 1) * (n[1] - 1);
    nPoints = nplane * (u64)(kn1 - kn0);
    nCells = eplane * (u64)(ke1 - ke0);
    p0 = nplane * (u64)kn0;
    c0 = eplane * (u64)ke0;
    nPointsG = nplane * (u64)n[2];
    nCellsG = eplane * (u64)(n[2] - 1);
    const std::string info = "TopOpt result version 1.1";
    const u64 off_pts = info.size() + 1 + 6 * sizeof(u64) + pnames.size() + 1 + cnames.size() + 1;
    const u64 off_conn = off_pts + 12 * nPointsG, off_off = off_conn + 64 * nCellsG, off_typ = off_off + 8 * nCellsG;
    data0 = off_typ + 8 * nCellsG;

    int fd = -1;
    if (rank == 0) {
        fd = open(filename.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd >= 0) {
            std::string h = info;
            h.push_back(0x01);
            u64 hdr[6] = {1, nPointsG, nCellsG, (u64)nPF, (u64)nCF, 8};
            h.append(reinterpret_cast<const char*>(hdr), sizeof hdr);
            h += pnames; h.push_back(0x01);
            h += cnames; h.push_back(0x01);
            if (!put(fd, h.data(), h.size(), 0)) { close(fd); fd = -1; }
        }
    }
    MPI_Barrier(PETSC_COMM_WORLD);  // the file exists (and is truncated) before anybody else touches it
    if (rank != 0) fd = open(filename.c_str(), O_WRONLY);
    if (fd < 0) { enabled = false; return; }

    const double dx = (xc[1] - xc[0]) / (n[0] - 1), dy = (xc[3] - xc[2]) / (n[1] - 1), dz = (xc[5] - xc[4]) / (n[2] - 1);
    std::vector<float> pts(3 * nPoints);
    size_t p = 0;
    for (int k = kn0; k < kn1; k++)
        for (int j = 0; j < n[1]; j++)
            for (int i = 0; i < n[0]; i++) {
                pts[p++] = (float)(xc[0] + i * dx); pts[p++] = (float)(xc[2] + j * dy); pts[p++] = (float)(xc[4] + k * dz);
            }
    bool ok = put(fd, pts.data(), pts.size() * sizeof(float), off_pts + 12 * p0);
    std::vector<u64> conn(8 * nCells), off(nCells), typ(nCells, 12);
    size_t e = 0;
    auto id = [&](int i, int j, int k) { return (u64)i + (u64)n[0] * (j + (u64)n[1] * k); };
    for (int k = ke0; k < ke1; k++)
        for (int j = 0; j < n[1] - 1; j++)
            for (int i = 0; i < n[0] - 1; i++) {
                u64* c = &conn[8 * e];
                c[0] = id(i, j, k); c[1] = id(i + 1, j, k); c[2] = id(i + 1, j + 1, k); c[3] = id(i, j + 1, k);
                c[4] = id(i, j, k + 1); c[5] = id(i + 1, j, k + 1); c[6] = id(i + 1, j + 1, k + 1); c[7] = id(i, j + 1, k + 1);
                off[e] = 8 * (c0 + e + 1);
                e++;
            }
    ok = ok && put(fd, conn.data(), conn.size() * sizeof(u64), off_conn + 64 * c0);
    ok = ok && put(fd, off.data(), off.size() * sizeof(u64), off_off + 8 * c0);
    ok = ok && put(fd, typ.data(), typ.size() * sizeof(u64), off_typ + 8 * c0);
    close(fd);
    if (!ok) enabled = false;
}
MPIIO::~MPIIO() {}

PetscErrorCode MPIIO::WriteVTK(DM, Vec U, Vec x, Vec xTilde, Vec xPhys, PetscInt itr) {
    if (!enabled) return 0;
    int fd = open(filename.c_str(), O_WRONLY);
    if (fd < 0) return 66;
    const u64 dump_bytes = 8 + 4 * ((u64)nPF * nPointsG + (u64)nCF * nCellsG);
    const u64 base = data0 + ndumps * dump_bytes;
    bool ok = true;
    if (TobRank() == 0) {
        u64 step = (u64)itr;
        ok = put(fd, &step, sizeof step, base);
    }
    PetscScalar *up, *xp, *xt, *xph;
    VecGetArray(U, &up); VecGetArray(x, &xp); VecGetArray(xTilde, &xt); VecGetArray(xPhys, &xph);
    std::vector<float> w(nPoints > nCells ? nPoints : nCells);
    for (int c = 0; c < nPF; c++) {  // ux, uy, uz (MPIIO.cc:169-178), fp32 down-cast
        for (u64 i = 0; i < nPoints; i++) w[i] = c < 3 ? (float)up[3 * i + c] : 0.0f;
        ok = ok && put(fd, w.data(), sizeof(float) * nPoints, base + 8 + 4 * ((u64)c * nPointsG + p0));
    }
    PetscScalar* cells[3] = {xp, xt, xph};  // x, xTilde, xPhys (MPIIO.cc:188-193)
    PetscInt nloc;
    VecGetLocalSize(x, &nloc);
    const u64 cbase = base + 8 + 4 * (u64)nPF * nPointsG;
    for (int c = 0; c < nCF; c++) {
        for (u64 i = 0; i < nCells; i++) w[i] = (c < 3 && i < (u64)nloc) ? (float)cells[c][i] : 0.0f;
        ok = ok && put(fd, w.data(), sizeof(float) * nCells, cbase + 4 * ((u64)c * nCellsG + c0));
    }
    close(fd);
    ndumps++;
    return ok ? 0 : 66;
}
