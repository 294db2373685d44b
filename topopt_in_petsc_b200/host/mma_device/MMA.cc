// MMA.cc (device flavour) -- forwards the MMA class surface to the tob_mma C-ABI.
#include "MMA.h"
#include <vector>
#include "../../../include/topopt_b200.h"

extern tob_ctx* TobDefaultCtx();  // set by TopOpt when it builds the grid

void MMA::Create(PetscInt m, PetscScalar* a, PetscScalar* c, PetscScalar* d) {
    m_ = m;
    if (tob_mma_create(TobDefaultCtx(), m, a, c, d, &h) != 0) {
        PetscErrorPrintf("tob_mma_create: %s\n", tob_last_error());
        MPI_Abort(PETSC_COMM_WORLD, 77);
    }
}
MMA::MMA(PetscInt, PetscInt m, Vec) : h(NULL) { Create(m, NULL, NULL, NULL); }
MMA::MMA(PetscInt, PetscInt m, Vec, PetscScalar* a, PetscScalar* c, PetscScalar* d) : h(NULL) { Create(m, a, c, d); }
MMA::MMA(PetscInt n, PetscInt m, PetscInt itr, Vec xo1, Vec xo2, Vec U, Vec L) : MMA(n, m, itr, xo1, xo2, U, L, NULL, NULL, NULL) {}
MMA::MMA(PetscInt, PetscInt m, PetscInt itr, Vec xo1, Vec xo2, Vec U, Vec L, PetscScalar* a, PetscScalar* c, PetscScalar* d) : h(NULL) {
    Create(m, a, c, d);
    if (itr < 3) PetscPrintf(PETSC_COMM_WORLD, "NOT A LEGAL RESTART POINT (k<3)\n");  // MMA.cc:27-29
    PetscScalar *p1, *p2, *pu, *pl;
    VecGetArray(xo1, &p1); VecGetArray(xo2, &p2); VecGetArray(U, &pu); VecGetArray(L, &pl);
    tob_mma_set_restart(h, p1, p2, pu, pl, itr);
}
MMA::~MMA() { if (h) tob_mma_destroy(h); }

PetscErrorCode MMA::Update(Vec xval, Vec dfdx, PetscScalar* gx, Vec* dgdx, Vec xmin, Vec xmax) {
    PetscScalar *x, *df, *lo, *hi;
    VecGetArray(xval, &x); VecGetArray(dfdx, &df); VecGetArray(xmin, &lo); VecGetArray(xmax, &hi);
    std::vector<double*> dg(m_);
    for (PetscInt j = 0; j < m_; j++) VecGetArray(dgdx[j], &dg[j]);
    int rc = tob_mma_update(h, x, df, gx, dg.data(), lo, hi);
    if (rc) PetscErrorPrintf("tob_mma_update: %s\n", tob_last_error());
    return rc;
}
PetscErrorCode MMA::Restart(Vec xo1, Vec xo2, Vec U, Vec L) {
    PetscScalar *p1, *p2, *pu, *pl;
    VecGetArray(xo1, &p1); VecGetArray(xo2, &p2); VecGetArray(U, &pu); VecGetArray(L, &pl);
    int k;
    return tob_mma_get_restart(h, p1, p2, pu, pl, &k);
}
PetscErrorCode MMA::SetOuterMovelimit(PetscScalar Xmin, PetscScalar Xmax, PetscScalar movelim, Vec x, Vec xmin, Vec xmax) {
    PetscScalar *xp, *lo, *hi;
    VecGetArray(x, &xp); VecGetArray(xmin, &lo); VecGetArray(xmax, &hi);
    return tob_mma_set_outer_movelimit(h, Xmin, Xmax, movelim, xp, lo, hi);
}
PetscScalar MMA::DesignChange(Vec x, Vec xold) {
    PetscScalar *xp, *xo;
    VecGetArray(x, &xp); VecGetArray(xold, &xo);
    double ch = 0;
    tob_mma_design_change(h, xp, xo, &ch);
    return ch;
}
