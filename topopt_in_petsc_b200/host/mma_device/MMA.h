// MMA.h (device flavour) -- the reference's MMA class surface (MMA.h:29-140) forwarding to tob_mma, the
// GPU implementation of the same algorithm (csrc/mma.cu).  Used by the product's own driver
// (host/topopt_b200_main.cc); the drop-in build instead compiles the reference's untouched MMA.cc/MMA.h.
#ifndef TOB_MMA_DEVICE_H
#define TOB_MMA_DEVICE_H
#include <petsc.h>
struct tob_mma;

class MMA {
  public:
    MMA(PetscInt n, PetscInt m, Vec x);
    MMA(PetscInt n, PetscInt m, Vec x, PetscScalar* a, PetscScalar* c, PetscScalar* d);
    MMA(PetscInt n, PetscInt m, PetscInt itr, Vec xo1, Vec xo2, Vec U, Vec L);
    MMA(PetscInt n, PetscInt m, PetscInt itr, Vec xo1, Vec xo2, Vec U, Vec L, PetscScalar* a, PetscScalar* c, PetscScalar* d);
    ~MMA();
    PetscErrorCode Update(Vec xval, Vec dfdx, PetscScalar* gx, Vec* dgdx, Vec xmin, Vec xmax);
    PetscErrorCode Restart(Vec xo1, Vec xo2, Vec U, Vec L);
    PetscErrorCode SetOuterMovelimit(PetscScalar Xmin, PetscScalar Xmax, PetscScalar movelim, Vec x, Vec xmin, Vec xmax);
    PetscScalar DesignChange(Vec x, Vec xold);

  private:
    tob_mma* h;
    PetscInt m_;
    void Create(PetscInt m, PetscScalar* a, PetscScalar* c, PetscScalar* d);
};
#endif
