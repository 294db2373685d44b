// TopOpt.h -- problem container with the reference's public surface (TopOpt.h:36-118), so that the
// untouched main.cc compiles against it.  The two DMDAs of the reference (TopOpt.cc:233-300) become one
// tob_ctx (grid + z-slab partition + stream + NCCL) shared by the nodal and element "DM" handles.
#ifndef TOB_TOPOPT_H
#define TOB_TOPOPT_H
#include <fstream>
#include <string>
#include <petsc.h>
#include "MMA.h"

class TopOpt {
  public:
    TopOpt(PetscInt nconstraints);
    TopOpt();
    ~TopOpt();

    PetscErrorCode AllocateMMAwithRestart(PetscInt* itr, MMA** mma);  // TopOpt.cc:386-512
    PetscErrorCode WriteRestartFiles(PetscInt* itr, MMA* mma);        // TopOpt.cc:514-570

    // physical domain (TopOpt.h:48-58)
    PetscScalar xc[6];
    PetscScalar dx, dy, dz;
    PetscInt nxyz[3];
    PetscInt nlvls;
    PetscScalar nu;
    DM da_nodes;
    DM da_elem;

    // optimisation parameters (TopOpt.h:60-84)
    PetscInt n, nloc, m;
    PetscScalar fx, fscale;
    PetscScalar* gx;
    PetscScalar Xmin, Xmax, movlim, volfrac, penal, Emin, Emax;
    PetscInt maxItr;
    PetscScalar rmin;
    PetscInt filter;
    PetscBool projectionFilter;
    PetscReal beta, betaFinal, eta;

    Vec x, xTilde, xPhys, dfdx, xmin, xmax, xold;
    Vec* dgdx;

    // MMA restart data (TopOpt.h:93-97)
    PetscBool restart, flip;
    std::string restdens_1, restdens_2;
    Vec xo1, xo2, U, L;

  private:
    void Defaults();
    PetscErrorCode BuildMesh();
    PetscErrorCode BuildDesignVectors();
    std::string file00, file00Itr, file01, file01Itr;
    static bool Exists(const std::string& f) { std::ifstream s(f.c_str()); return (bool)s; }
};
#endif
