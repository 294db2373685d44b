// galerkin.cuh -- assembled Galerkin coarse levels (27-point 3x3-block stencils) of the multigrid preconditioner
#pragma once
#include "common.cuh"

namespace tob {

struct StencilLevel {
    void* A = nullptr;       // [243][nnp] float (default) or double: k = 9 * offset + 3 * c + c2, offset = ((dk+1)*3 + (dj+1))*3 + (di+1)
    bool fp32 = true;
    long long nnp = 0;       // nzn * nyn * px nodes (x padded like the nodal vectors)
    double* coef = nullptr;  // Chebyshev {s0, s1} pairs (device; refreshed with the operator)
    double* lam = nullptr;   // power-iteration estimate of lambda_max(D^-1 A)
    int* its = nullptr;      // iterations of the last single-CTA coarse solve
    double *pv = nullptr, *pt = nullptr, *pw = nullptr;  // power-iteration work vectors
    bool pv_init = false;
    double* Ainv = nullptr;  // coarsest level: explicit inverse (dense, ndense x ndense), refreshed with the operator
    int ndense = 0;
    double* prow = nullptr;      // double-buffered pivot row of the inversion
    unsigned int* bar = nullptr; // its grid-barrier counter
    double* M = nullptr;     // [8][24][24] Q_child^T KE Q_child (level assembled from a matrix-free parent)
};

constexpr int kDenseCoarseMax = 1024;  // dofs up to which the coarsest level is inverted explicitly
int stencil_dense_setup(tob_ctx_impl* c, const LevelGrid& g, StencilLevel& S);
int stencil_dense_solve(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, const double* b, double* x);
void galerkin_child_matrices(const double* KE_fine, double* M);
int stencil_alloc(tob_ctx_impl* c, const LevelGrid& g, StencilLevel& S);
void stencil_free(StencilLevel& S);
int stencil_assemble_from_mf(tob_ctx_impl* c, const LevelGrid& gf, const double* Ef, const double* KE_fine, const LevelGrid& gc,
                             const unsigned char* maskc, StencilLevel& S);
int stencil_assemble_from_stencil(tob_ctx_impl* c, const LevelGrid& gf, const StencilLevel& F, const LevelGrid& gc, const unsigned char* maskc,
                                  StencilLevel& S);
int stencil_finish(tob_ctx_impl* c, const LevelGrid& g, StencilLevel& S, double* dinv, double ratio, int deg, int nit);
int stencil_apply(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, int epi, const double* x, double* y, const double* a0,
                  const double* a1, double* o1, int step, int* nparts);
int stencil_cheb_first(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, double* x, double* d, const double* b, const double* dinv);
int stencil_coarse_solve(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, const double* dinv, const double* b, double* x, double* r,
                         double* p, double* q, double rtol, int maxit);

}  // namespace tob
