/* petsc.h -- the slice of the PETSc API that the reference's main.cc and MMA.cc touch,
 * implemented over pinned host arrays.  Vec is a host-resident array (VecGetArray must hand back CPU
 * pointers: MMA.cc and main.cc dereference them); the B200 classes pass those pointers straight into the
 * C-ABI of libtopopt_b200.so, which stages them to the GPU.  DM is an opaque handle carrying the grid and
 * the tob_ctx.  Not a PETSc re-implementation: only what the untouched callers need. */
#ifndef TOB_PETSC_SHIM_H
#define TOB_PETSC_SHIM_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "mpi.h"

typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
#define PETSC_NULL NULL
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_COMM_SELF MPI_COMM_WORLD
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_DECIDE (-1)
#define PETSC_DEFAULT (-2)
#define MPIU_SCALAR MPI_DOUBLE
#define MPIU_REAL MPI_DOUBLE
#define MPIU_INT MPI_INT
#define CHKERRQ(ierr) do { if ((ierr) != 0) return (ierr); } while (0)
#define PetscAbsReal(a) fabs(a)
#define PetscAbsScalar(a) fabs(a)
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscPowScalar(a, b) pow(a, b)
#define PetscSqrtScalar(a) sqrt(a)

struct tob_ctx;
struct _p_Vec;
struct _p_DM;
struct _p_PetscViewer;
typedef struct _p_Vec* Vec;
typedef struct _p_DM* DM;
typedef struct _p_PetscViewer* PetscViewer;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_INFINITY = 3 } NormType;
typedef enum { FILE_MODE_READ = 0, FILE_MODE_WRITE = 1 } PetscFileMode;
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;

PetscErrorCode PetscInitialize(int* argc, char*** argv, const char* file, const char* help);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char* fmt, ...);
PetscErrorCode PetscErrorPrintf(const char* fmt, ...);
PetscErrorCode PetscOptionsGetInt(void* opts, const char* pre, const char* name, PetscInt* v, PetscBool* set);
PetscErrorCode PetscOptionsGetReal(void* opts, const char* pre, const char* name, PetscReal* v, PetscBool* set);
PetscErrorCode PetscOptionsGetBool(void* opts, const char* pre, const char* name, PetscBool* v, PetscBool* set);
PetscErrorCode PetscOptionsGetString(void* opts, const char* pre, const char* name, char* v, size_t len, PetscBool* set);

/* ---- Vec --------------------------------------------------------------------------------------- */
PetscErrorCode VecCreateTob(PetscInt nlocal, Vec* v); /* shim-internal constructor (global size = sum over ranks) */
PetscErrorCode VecDuplicate(Vec v, Vec* out);
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt m, Vec** out);
PetscErrorCode VecDestroy(Vec* v);
PetscErrorCode VecDestroyVecs(PetscInt m, Vec** v);
PetscErrorCode VecGetArray(Vec v, PetscScalar** a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar** a);
PetscErrorCode VecGetArrays(const Vec* v, PetscInt m, PetscScalar*** a);
PetscErrorCode VecRestoreArrays(const Vec* v, PetscInt m, PetscScalar*** a);
PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n);
PetscErrorCode VecGetSize(Vec v, PetscInt* n);
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecScale(Vec v, PetscScalar a);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar a, PetscScalar b, PetscScalar g, Vec x, Vec y);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);
PetscErrorCode VecSum(Vec v, PetscScalar* s);
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal* n);
PetscErrorCode VecView(Vec v, PetscViewer viewer);
PetscErrorCode VecLoad(Vec v, PetscViewer viewer);

/* ---- viewers: PETSc binary Vec stream (restart files, TopOpt.cc:514-570) and a tiny ASCII printer ----- */
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm comm, const char* name, PetscFileMode mode, PetscViewer* v);
PetscErrorCode PetscViewerASCIIOpen(MPI_Comm comm, const char* name, PetscViewer* v);
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char* fmt, ...);
PetscErrorCode PetscViewerDestroy(PetscViewer* v);

/* ---- DM: grid handle ------------------------------------------------------------------------------ */
PetscErrorCode DMCreateTob(PetscInt nxn, PetscInt nyn, PetscInt nzn, const PetscScalar xc[6], PetscInt dof, DM* dm);
PetscErrorCode DMCloneTob(DM dm, PetscInt dof, DM* out); /* shares the tob_ctx (reference counted) */
PetscErrorCode DMDestroy(DM* dm);
struct tob_ctx* DMGetTobCtx(DM dm);
PetscErrorCode DMGetGridTob(DM dm, PetscInt n[3], PetscScalar xc[6]);
int TobRank(void);
int TobSize(void);

#endif
