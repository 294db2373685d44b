// shim_selftest.cc -- CPU-only self test of the host shim (petsc.h / mpi.h stand-ins): Vec algebra, option parsing,
// PETSc-binary viewer round trip, and -- when started as 2 ranks (TOB_RANK/TOB_NRANKS/TOB_JOB) -- the shared-memory
// mini-MPI: Allreduce, Allgather, parallel VecView + VecLoad.  With -DWITH_REF_MMA (built by oracle/Makefile where the
// reference checkout exists) it also runs the reference's unmodified MMA.cc split over the ranks and prints the new
// design's checksum, which must not depend on the number of ranks.
#include <petsc.h>
#include <vector>
#ifdef WITH_REF_MMA
#include "MMA.h"
#endif

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "[rank %d] FAILED: %s (line %d)\n", TobRank(), #c, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    PetscInitialize(&argc, &argv, NULL, NULL);
    const int rank = TobRank(), size = TobSize();
    // options
    PetscInt n = 0; PetscReal r = 0; PetscBool b = PETSC_TRUE, flg; char s[64];
    PetscOptionsGetInt(NULL, NULL, "-n", &n, &flg); REQUIRE(flg && n == 1000);
    PetscOptionsGetReal(NULL, NULL, "-neg", &r, &flg); REQUIRE(flg && r == -2.5);
    PetscOptionsGetBool(NULL, NULL, "-off", &b, &flg); REQUIRE(flg && b == PETSC_FALSE);
    PetscOptionsGetString(NULL, NULL, "-name", s, sizeof s, &flg); REQUIRE(flg && std::string(s) == "abc");
    PetscOptionsGetInt(NULL, NULL, "-missing", &n, &flg); REQUIRE(!flg && n == 1000);
    // Vec algebra on a vector split unevenly over the ranks
    const PetscInt nglob = n, nloc = rank == 0 ? nglob - (size - 1) * (nglob / size - 3) : nglob / size - 3;
    PetscInt off = 0;
    { std::vector<int> cnt(size); int me = nloc; MPI_Allgather(&me, 1, MPI_INT, cnt.data(), 1, MPI_INT, MPI_COMM_WORLD); for (int k = 0; k < rank; k++) off += cnt[k]; }
    Vec x, y, z;
    VecCreateTob(nloc, &x); VecDuplicate(x, &y); VecDuplicate(x, &z);
    PetscInt ng; VecGetSize(x, &ng); REQUIRE(ng == nglob);
    PetscScalar *xp, *yp;
    VecGetArray(x, &xp); VecGetArray(y, &yp);
    for (PetscInt i = 0; i < nloc; i++) { xp[i] = 0.001 * (off + i); yp[i] = 1.0; }
    PetscScalar sum; VecSum(x, &sum); REQUIRE(fabs(sum - 0.001 * (double)nglob * (nglob - 1) / 2) < 1e-9);
    PetscReal nrm; VecNorm(y, NORM_2, &nrm); REQUIRE(fabs(nrm - sqrt((double)nglob)) < 1e-12);
    VecNorm(x, NORM_INFINITY, &nrm); REQUIRE(fabs(nrm - 0.001 * (nglob - 1)) < 1e-15);
    VecAXPBYPCZ(z, 2.0, -1.0, 0.0, x, y); VecAXPY(z, 0.5, y); VecScale(z, 2.0);  // z = 2(2x - y + 0.5y) = 4x - y
    VecSum(z, &sum); REQUIRE(fabs(sum - (4 * 0.001 * (double)nglob * (nglob - 1) / 2 - nglob)) < 1e-8);
    // mini-MPI reductions
    double v[3] = {1.0 * (rank + 1), -1.0 * rank, 3.0}, w[3];
    MPI_Allreduce(v, w, 3, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    REQUIRE(w[0] == size * (size + 1) / 2.0 && w[2] == 3.0 * size);
    MPI_Allreduce(v, w, 3, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    REQUIRE(w[0] == size && w[1] == 0.0);
    // PETSc binary Vec stream: parallel write, parallel read back into another vector
    const char* fn = "shim_selftest_vec.dat";
    PetscViewer vw;
    REQUIRE(PetscViewerBinaryOpen(PETSC_COMM_WORLD, fn, FILE_MODE_WRITE, &vw) == 0);
    VecView(x, vw); VecView(z, vw);
    PetscViewerDestroy(&vw);
    MPI_Barrier(MPI_COMM_WORLD);
    Vec x2, z2; VecDuplicate(x, &x2); VecDuplicate(x, &z2);
    REQUIRE(PetscViewerBinaryOpen(PETSC_COMM_WORLD, fn, FILE_MODE_READ, &vw) == 0);
    REQUIRE(VecLoad(x2, vw) == 0 && VecLoad(z2, vw) == 0);
    PetscViewerDestroy(&vw);
    VecAXPY(x2, -1.0, x); VecAXPY(z2, -1.0, z);
    VecNorm(x2, NORM_INFINITY, &nrm); REQUIRE(nrm == 0.0);
    VecNorm(z2, NORM_INFINITY, &nrm); REQUIRE(nrm == 0.0);
    if (rank == 0) {  // header check: big-endian class id 1211214 and length
        FILE* f = fopen(fn, "rb"); unsigned char h[8]; REQUIRE(fread(h, 1, 8, f) == 8); fclose(f);
        REQUIRE(((h[0] << 24) | (h[1] << 16) | (h[2] << 8) | h[3]) == 1211214);
        REQUIRE(((h[4] << 24) | (h[5] << 16) | (h[6] << 8) | h[7]) == nglob);
    }
    MPI_Barrier(MPI_COMM_WORLD);
    if (rank == 0) remove(fn);
#ifdef WITH_REF_MMA
    {   // the reference's own MMA on a synthetic problem, distributed over the ranks
        Vec xm, df, lo, hi, xold, *dg;
        VecDuplicate(x, &xm); VecDuplicate(x, &df); VecDuplicate(x, &lo); VecDuplicate(x, &hi); VecDuplicate(x, &xold);
        VecDuplicateVecs(x, 1, &dg);
        VecSet(xm, 0.3); VecSet(xold, 0.3); VecSet(dg[0], 1.0 / nglob);
        MMA mma(nglob, 1, xm);
        PetscScalar *xa, *dfa;
        double checksum = 0;
        for (int it = 0; it < 5; it++) {
            VecGetArray(xm, &xa); VecGetArray(df, &dfa);
            for (PetscInt i = 0; i < nloc; i++) { const double c0 = 0.5 + 1.5 * fabs(sin(0.37 * (off + i) + it)); dfa[i] = -c0 / ((xa[i] + 0.1) * (xa[i] + 0.1)); }
            PetscScalar vol; VecSum(xm, &vol);
            PetscScalar gx[1] = {vol / nglob - 0.3};
            mma.SetOuterMovelimit(0.0, 1.0, 0.2, xm, lo, hi);
            mma.Update(xm, df, gx, dg, lo, hi);
            mma.DesignChange(xm, xold);
            VecGetArray(xm, &xa);
            double loc = 0, tot = 0;
            for (PetscInt i = 0; i < nloc; i++) loc += xa[i] * (1.0 + 0.001 * ((off + i) % 97));
            MPI_Allreduce(&loc, &tot, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
            checksum = tot;
        }
        if (rank == 0) printf("MMA_CHECKSUM %.12e\n", checksum);
    }
#endif
    if (rank == 0) printf("SHIM_SELFTEST_OK ranks=%d\n", size);
    PetscFinalize();
    return 0;
}
