/* petsc/private/dmdaimpl.h (petsc_lite) -- the reference reaches into DM_DA for its element cache
 * (DMDAGetElements_3D, LinearElasticity.cc:785-839): petsc.h of petsc_lite defines the struct. */
#include <petsc.h>
