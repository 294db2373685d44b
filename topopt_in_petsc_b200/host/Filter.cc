// Filter.cc -- forwarding wrapper for Filter and PDEFilt.
#include "Filter.h"
#include <vector>
#include "../../include/topopt_b200.h"

static void die_on(int rc, const char* what) {
    if (rc != 0) {
        PetscErrorPrintf("%s: %s\n", what, tob_last_error());
        MPI_Abort(PETSC_COMM_WORLD, 77);
    }
}

Filter::Filter(DM da_nodes, Vec, PetscInt filterT, PetscScalar Rin) : h(NULL), filterType(filterT), R(Rin) {
    die_on(tob_filter_create(DMGetTobCtx(da_nodes), filterT, Rin, &h), "tob_filter_create");
    if (filterT == 0 || filterT == 1) {
        int s = 0;
        tob_filter_elemconn(h, &s);
        PetscPrintf(PETSC_COMM_WORLD, "# Filter radius rmin = %f results in a stencil of %i elements \n", R, s);
    }
}
Filter::~Filter() { if (h) tob_filter_destroy(h); }

PetscErrorCode Filter::FilterProject(Vec x, Vec xTilde, Vec xPhys, PetscBool projectionFilter, PetscScalar beta, PetscScalar eta) {
    PetscScalar *xp, *xt, *xph;
    VecGetArray(x, &xp); VecGetArray(xTilde, &xt); VecGetArray(xPhys, &xph);
    int rc = tob_filter_project(h, xp, xt, xph, projectionFilter ? 1 : 0, beta, eta);
    if (rc) PetscErrorPrintf("tob_filter_project: %s\n", tob_last_error());
    return rc;
}
PetscErrorCode Filter::Gradients(Vec x, Vec xTilde, Vec dfdx, PetscInt m, Vec* dgdx, PetscBool projectionFilter, PetscScalar beta,
                                 PetscScalar eta) {
    PetscScalar *xp, *xt, *df;
    VecGetArray(x, &xp); VecGetArray(xTilde, &xt); VecGetArray(dfdx, &df);
    std::vector<double*> dg(m);
    for (PetscInt i = 0; i < m; i++) VecGetArray(dgdx[i], &dg[i]);
    int rc = tob_filter_gradients(h, xp, xt, df, m, dg.data(), projectionFilter ? 1 : 0, beta, eta);
    if (rc) PetscErrorPrintf("tob_filter_gradients: %s\n", tob_last_error());
    return rc;
}
PetscBool Filter::IncreaseBeta(PetscReal* beta, PetscReal betaFinal, PetscScalar gx, PetscInt itr, PetscReal ch) {
    int changed = 0;
    const PetscReal before = *beta;
    tob_filter_increase_beta(beta, betaFinal, gx, itr, ch, &changed);
    if (*beta != before) PetscPrintf(PETSC_COMM_WORLD, "Beta has been increased to: %f\n", beta[0]);  // Filter.cc:284
    return changed ? PETSC_TRUE : PETSC_FALSE;
}
PetscScalar Filter::GetMND(Vec x) {
    PetscScalar* xp;
    VecGetArray(x, &xp);
    double v = 0;
    die_on(tob_filter_mnd(h, xp, &v), "tob_filter_mnd");
    return v;
}

PDEFilt::PDEFilt(DM da_nodes, PetscScalar rmin) : h(NULL) { die_on(tob_filter_create(DMGetTobCtx(da_nodes), 2, rmin, &h), "tob_filter_create"); }
PDEFilt::~PDEFilt() { if (h) tob_filter_destroy(h); }
PetscErrorCode PDEFilt::FilterProject(Vec XX, Vec F) {
    PetscScalar *a, *b;
    VecGetArray(XX, &a); VecGetArray(F, &b);
    return tob_filter_pde_apply(h, a, b);
}
PetscErrorCode PDEFilt::Gradients(Vec OS, Vec FS) { return FilterProject(OS, FS); }
