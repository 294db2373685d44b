// PDEFilter.h -- class surface of the reference's PDEFilt (PDEFilter.h:29-74).  The Helmholtz solve lives
// inside tob_filter (filter type 2); this wrapper exists so code that names PDEFilt keeps compiling.
#ifndef TOB_PDEFILTER_H
#define TOB_PDEFILTER_H
#include <petsc.h>
struct tob_filter;
class PDEFilt {
  public:
    PDEFilt(DM da_nodes, PetscScalar rmin);
    ~PDEFilt();
    PetscErrorCode FilterProject(Vec XX, Vec F);
    PetscErrorCode Gradients(Vec OS, Vec FS);
  private:
    tob_filter* h;
};
#endif
