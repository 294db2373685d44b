// MPIIO.cc -- writes the reference's binary result container with plain fwrite (one writer).
// Layout (MPIIO.cc:293-371 header, :380-508 mesh, :511-714 fields; reader bin2vtu_v3.py:23-144):
//   info string + 0x01 | u64 nDom | u64 nPointsT[nDom] nCellsT[] nPFields[] nCFields[] | u64 nodesPerElement
//   | point-field names + 0x01 | cell-field names + 0x01 | f32 xyz per point | u64 connectivity (8/cell)
//   | u64 offsets | u64 cell types (12 = VTK hexahedron) | per dump: u64 iteration, f32 point fields
//   (field-major), f32 cell fields (field-major).
// Multi-rank runs: only rank 0's slab is written in this round (documented limitation).
#include "MPIIO.h"
#include <vector>
#include "../../include/topopt_b200.h"

static void put_string(FILE* f, const std::string& s) { fwrite(s.data(), 1, s.size(), f); fputc(0x01, f); }

MPIIO::MPIIO(DM da_nodes, int nPfields, std::string pnames, int nCfields, std::string cnames)
    : nPoints(0), nCells(0), nPF(nPfields), nCF(nCfields), enabled(TobRank() == 0) {
    PetscBool flg;
    char buf[PETSC_MAX_PATH_LEN];
    filename = "output_00000.dat";
    PetscOptionsGetString(NULL, NULL, "-workdir", buf, sizeof(buf), &flg);
    if (flg) filename = std::string(buf) + "/output.dat";  // MPIIO.cc:229-236
    PetscBool on = PETSC_TRUE;
    PetscOptionsGetBool(NULL, NULL, "-tob_output", &on, &flg);
    if (!on) enabled = false;
    if (!enabled) return;
    PetscInt n[3]; PetscScalar xc[6];
    DMGetGridTob(da_nodes, n, xc);
    int kn0, kn1, ke0, ke1;
    tob_ctx_slab(DMGetTobCtx(da_nodes), &kn0, &kn1, &ke0, &ke1);
    const int nzp = kn1 - kn0;
    nPoints = (unsigned long)n[0] * n[1] * nzp;
    nCells = (unsigned long)(n[0] - 1) * (n[1] - 1) * (nzp - 1);
    FILE* f = fopen(filename.c_str(), "wb");
    if (!f) { enabled = false; return; }
    put_string(f, "TopOpt result version 1.1");
    unsigned long hdr[6] = {1, nPoints, nCells, (unsigned long)nPF, (unsigned long)nCF, 8};
    fwrite(hdr, sizeof(unsigned long), 6, f);
    put_string(f, pnames);
    put_string(f, cnames);
    const double dx = (xc[1] - xc[0]) / (n[0] - 1), dy = (xc[3] - xc[2]) / (n[1] - 1), dz = (xc[5] - xc[4]) / (n[2] - 1);
    std::vector<float> pts(3 * nPoints);
    size_t p = 0;
    for (int k = 0; k < nzp; k++)
        for (int j = 0; j < n[1]; j++)
            for (int i = 0; i < n[0]; i++) {
                pts[p++] = (float)(xc[0] + i * dx); pts[p++] = (float)(xc[2] + j * dy); pts[p++] = (float)(xc[4] + (k + kn0) * dz);
            }
    fwrite(pts.data(), sizeof(float), pts.size(), f);
    std::vector<unsigned long> conn(8 * nCells), off(nCells), typ(nCells, 12);
    size_t e = 0;
    auto id = [&](int i, int j, int k) { return (unsigned long)i + (unsigned long)n[0] * (j + (unsigned long)n[1] * k); };
    for (int k = 0; k < nzp - 1; k++)
        for (int j = 0; j < n[1] - 1; j++)
            for (int i = 0; i < n[0] - 1; i++) {
                unsigned long* c = &conn[8 * e];
                c[0] = id(i, j, k); c[1] = id(i + 1, j, k); c[2] = id(i + 1, j + 1, k); c[3] = id(i, j + 1, k);
                c[4] = id(i, j, k + 1); c[5] = id(i + 1, j, k + 1); c[6] = id(i + 1, j + 1, k + 1); c[7] = id(i, j + 1, k + 1);
                off[e] = 8 * (e + 1);
                e++;
            }
    fwrite(conn.data(), sizeof(unsigned long), conn.size(), f);
    fwrite(off.data(), sizeof(unsigned long), off.size(), f);
    fwrite(typ.data(), sizeof(unsigned long), typ.size(), f);
    fclose(f);
}
MPIIO::~MPIIO() {}

PetscErrorCode MPIIO::WriteVTK(DM, Vec U, Vec x, Vec xTilde, Vec xPhys, PetscInt itr) {
    if (!enabled) return 0;
    FILE* f = fopen(filename.c_str(), "ab");
    if (!f) return 66;
    unsigned long step = (unsigned long)itr;
    fwrite(&step, sizeof step, 1, f);
    PetscScalar *up, *xp, *xt, *xph;
    VecGetArray(U, &up); VecGetArray(x, &xp); VecGetArray(xTilde, &xt); VecGetArray(xPhys, &xph);
    std::vector<float> w(nPoints > nCells ? nPoints : nCells);
    for (int c = 0; c < nPF && c < 3; c++) {  // ux, uy, uz (MPIIO.cc:169-178), fp32 down-cast
        for (unsigned long i = 0; i < nPoints; i++) w[i] = (float)up[3 * i + c];
        fwrite(w.data(), sizeof(float), nPoints, f);
    }
    PetscScalar* cells[3] = {xp, xt, xph};  // x, xTilde, xPhys (MPIIO.cc:188-193)
    PetscInt nloc;
    VecGetLocalSize(x, &nloc);
    for (int c = 0; c < nCF && c < 3; c++) {
        for (unsigned long i = 0; i < nCells; i++) w[i] = i < (unsigned long)nloc ? (float)cells[c][i] : 0.0f;
        fwrite(w.data(), sizeof(float), nCells, f);
    }
    fclose(f);
    return 0;
}
