// LinearElasticity.h -- the reference's physics class surface (LinearElasticity.h:21-109) forwarding to
// the tob_elast C-ABI (matrix-free hex8 SpMV + PCG/GMG on the GPU).  No Mat, no KSP.
#ifndef TOB_LINEARELASTICITY_H
#define TOB_LINEARELASTICITY_H
#include <string>
#include <petsc.h>
struct tob_elast;

class LinearElasticity {
  public:
    LinearElasticity(DM da_nodes);
    ~LinearElasticity();
    PetscErrorCode ComputeObjectiveConstraintsSensitivities(PetscScalar* fx, PetscScalar* gx, Vec dfdx, Vec dgdx, Vec xPhys,
                                                            PetscScalar Emin, PetscScalar Emax, PetscScalar penal,
                                                            PetscScalar volfrac);
    PetscErrorCode ComputeObjectiveConstraints(PetscScalar* fx, PetscScalar* gx, Vec xPhys, PetscScalar Emin, PetscScalar Emax,
                                               PetscScalar penal, PetscScalar volfrac);
    PetscErrorCode ComputeSensitivities(Vec dfdx, Vec dgdx, Vec xPhys, PetscScalar Emin, PetscScalar Emax, PetscScalar penal,
                                        PetscScalar volfrac);
    PetscErrorCode WriteRestartFiles();
    Vec GetStateField();  // host copy of U, refreshed from the device
    DM GetDM() { return da_nodal; }
    DM da_nodal;

  private:
    tob_elast* h;
    Vec U;
    PetscInt nlvls;
    PetscScalar nu;
    PetscBool restart, flip;
    std::string file00, file01;
    bool state_loaded;
    void LoadRestartState();
    PetscErrorCode Report(int rc, bool solved);
};
#endif
