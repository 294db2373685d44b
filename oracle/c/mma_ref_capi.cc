// mma_ref_capi.cc -- thin extern "C" harness around the REFERENCE's own, unmodified MMA class
// (/root/reference/MMA.cc + MMA.h, compiled where they lie against the stub petsc.h).  TEST
// INFRASTRUCTURE ONLY: lets tests/test_mma_ref.py pin oracle/topopt_oracle.py:MMAOracle and the
// device MMA (csrc/mma.cu) against the reference's real optimiser code.
#include "MMA.h"
#include <vector>

struct MmaRef {
    int n, m;
    MMA* mma;
    Vec x, dfdx, xmin, xmax, xold;
    Vec* dgdx;
};

static void put(Vec v, const double* src, int n) {
    PetscScalar* p;
    VecGetArray(v, &p);
    memcpy(p, src, sizeof(double) * n);
    VecRestoreArray(v, &p);
}
static void get(Vec v, double* dst, int n) {
    PetscScalar* p;
    VecGetArray(v, &p);
    memcpy(dst, p, sizeof(double) * n);
    VecRestoreArray(v, &p);
}

extern "C" {
void* mmaref_create(int n, int m) {
    static bool init = false;
    if (!init) { PetscInitialize(nullptr, nullptr, nullptr, nullptr); init = true; }
    MmaRef* h = new MmaRef();
    h->n = n; h->m = m;
    VecCreateTob(n, &h->x);
    VecDuplicate(h->x, &h->dfdx);
    VecDuplicate(h->x, &h->xmin);
    VecDuplicate(h->x, &h->xmax);
    VecDuplicate(h->x, &h->xold);
    VecDuplicateVecs(h->x, m, &h->dgdx);
    h->mma = new MMA(n, m, h->x);  // MMA.cc:253-291
    return h;
}
void mmaref_destroy(void* hh) {
    MmaRef* h = (MmaRef*)hh;
    delete h->mma;
    VecDestroy(&h->x); VecDestroy(&h->dfdx); VecDestroy(&h->xmin); VecDestroy(&h->xmax); VecDestroy(&h->xold);
    VecDestroyVecs(h->m, &h->dgdx);
    delete h;
}
void mmaref_set_outer_movelimit(void* hh, double Xmin, double Xmax, double movlim, const double* x, double* xmin, double* xmax) {
    MmaRef* h = (MmaRef*)hh;
    put(h->x, x, h->n);
    h->mma->SetOuterMovelimit(Xmin, Xmax, movlim, h->x, h->xmin, h->xmax);
    get(h->xmin, xmin, h->n);
    get(h->xmax, xmax, h->n);
}
void mmaref_update(void* hh, double* x, const double* dfdx, const double* gx, const double* dgdx, const double* xmin, const double* xmax) {
    MmaRef* h = (MmaRef*)hh;
    put(h->x, x, h->n);
    put(h->dfdx, dfdx, h->n);
    put(h->xmin, xmin, h->n);
    put(h->xmax, xmax, h->n);
    for (int j = 0; j < h->m; j++) put(h->dgdx[j], dgdx + (size_t)j * h->n, h->n);
    std::vector<double> g(gx, gx + h->m);
    h->mma->Update(h->x, h->dfdx, g.data(), h->dgdx, h->xmin, h->xmax);
    get(h->x, x, h->n);
}
double mmaref_design_change(void* hh, const double* x, double* xold) {
    MmaRef* h = (MmaRef*)hh;
    put(h->x, x, h->n);
    put(h->xold, xold, h->n);
    double ch = h->mma->DesignChange(h->x, h->xold);
    get(h->xold, xold, h->n);
    return ch;
}
}
