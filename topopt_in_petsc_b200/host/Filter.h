// Filter.h -- the reference's Filter class surface (Filter.h:34-92) over the tob_filter C-ABI
// (stencil cone filter / Helmholtz PDE filter / Heaviside projection kernels on the GPU).
#ifndef TOB_FILTER_H
#define TOB_FILTER_H
#include <petsc.h>
#include "PDEFilter.h"
struct tob_filter;

class Filter {
  public:
    Filter(DM da_nodes, Vec x, PetscInt filterT, PetscScalar Rin);
    ~Filter();
    PetscErrorCode FilterProject(Vec x, Vec xTilde, Vec xPhys, PetscBool projectionFilter, PetscScalar beta, PetscScalar eta);
    PetscErrorCode Gradients(Vec x, Vec xTilde, Vec dfdx, PetscInt m, Vec* dgdx, PetscBool projectionFilter, PetscScalar beta,
                             PetscScalar eta);
    PetscBool IncreaseBeta(PetscReal* beta, PetscReal betaFinal, PetscScalar gx, PetscInt itr, PetscReal ch);
    PetscScalar GetMND(Vec x);

  private:
    tob_filter* h;
    PetscInt filterType;
    PetscScalar R;
};
#endif
