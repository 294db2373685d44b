/* petsc.h (petsc_lite) -- a SERIAL, CPU-only restatement of the slice of the PETSc API the reference uses, written so
 * that the reference's own, unmodified sources (main.cc, TopOpt.cc, LinearElasticity.cc, Filter.cc, PDEFilter.cc,
 * MMA.cc, MPIIO.cc) compile and run on one process.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): it exists to pin the numpy/C oracles and the CUDA product against outputs of the
 * reference itself.  PETSc is an un-vendored dependency of the reference; what is restated here is its
 * documented behaviour for: Vec (dense arrays), Mat (AIJ with local == global numbering), DMDA (one process, no
 * ghost points), binary/ASCII viewers, the options database, and KSP.  KSPSolve is the one deliberate deviation: the
 * reference's FGMRES+PCMG stack lives entirely inside PETSc, so petsc_lite solves the same (SPD) system with
 * Jacobi-PCG to a residual of 1e-13 ||b|| instead -- the comparison targets (KE, BC, assembled operator, compliance,
 * sensitivities, filters, MMA) do not depend on which Krylov method produced U.  PC/PCMG calls are recorded no-ops. */
#ifndef PETSC_LITE_H
#define PETSC_LITE_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpi.h"

typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
#define PETSC_NULL NULL
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_COMM_SELF MPI_COMM_SELF
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
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
#define PetscMalloc(sz, pp) (*(void**)(pp) = malloc(sz), 0)
#define PetscFree(p) (free((void*)(p)), (p) = NULL, 0)

typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_INFINITY = 3 } NormType;
typedef enum { FILE_MODE_READ = 0, FILE_MODE_WRITE = 1 } PetscFileMode;
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { DM_BOUNDARY_NONE = 0, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_MIRROR, DM_BOUNDARY_PERIODIC } DMBoundaryType;
typedef enum { DMDA_STENCIL_STAR = 0, DMDA_STENCIL_BOX } DMDAStencilType;
typedef enum { DMDA_ELEMENT_P1 = 0, DMDA_ELEMENT_Q1 } DMDAElementType;
typedef enum { PC_MG_MULTIPLICATIVE = 0, PC_MG_ADDITIVE, PC_MG_FULL, PC_MG_KASKADE } PCMGType;
typedef enum { PC_MG_CYCLE_V = 1, PC_MG_CYCLE_W = 2 } PCMGCycleType;
typedef enum { PC_MG_GALERKIN_BOTH = 0, PC_MG_GALERKIN_PMAT, PC_MG_GALERKIN_MAT, PC_MG_GALERKIN_NONE } PCMGGalerkinType;
typedef const char* KSPType;
typedef const char* PCType;
typedef const char* PetscViewerType;
typedef const char* PetscRandomType;
#define KSPCG "cg"
#define KSPGMRES "gmres"
#define KSPFGMRES "fgmres"
#define KSPCHEBYSHEV "chebyshev"
#define PCMG "mg"
#define PCSOR "sor"
#define PCJACOBI "jacobi"
#define PCGAMG "gamg"
#define PCSPAI "spai"
#define PETSCVIEWERASCII "ascii"
#define PETSCVIEWERBINARY "binary"
#define PETSCRAND48 "rand48"

struct _p_Vec;
struct _p_Mat;
struct _p_KSP;
struct _p_PC;
struct _p_PetscViewer;
struct _p_PetscRandom;
struct _p_ISL2G;
typedef struct _p_Vec* Vec;
typedef struct _p_Mat* Mat;
typedef struct _p_KSP* KSP;
typedef struct _p_PC* PC;
typedef struct _p_PetscViewer* PetscViewer;
typedef struct _p_PetscRandom* PetscRandom;
typedef struct _p_ISL2G* ISLocalToGlobalMapping;
typedef void* PetscObject;

/* the reference dereferences dm->data as DM_DA* and uses e / ne / elementtype (LinearElasticity.cc:787-833) */
typedef struct {
    PetscInt M, N, P;            /* grid points per direction */
    PetscInt dof, s;             /* dofs per point, stencil width */
    DMBoundaryType bx, by, bz;
    DMDAStencilType stencil;
    DMDAElementType elementtype;
    PetscInt* e;                 /* element cache filled by the caller's DMDAGetElements_3D */
    PetscInt ne;
    PetscInt lx[1], ly[1], lz[1];
    Vec coords;                  /* 3 coordinates per grid point */
} DM_DA;
struct _p_DM {
    void* data;
};
typedef struct _p_DM* DM;
typedef struct {
    PetscInt dim, dof, sw;
    PetscInt mx, my, mz;
    PetscInt xs, ys, zs, xm, ym, zm;
    PetscInt gxs, gys, gzs, gxm, gym, gzm;
    DMBoundaryType bx, by, bz;
    DMDAStencilType st;
    DM da;
} DMDALocalInfo;

PetscErrorCode PetscInitialize(int* argc, char*** argv, const char* file, const char* help);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char* fmt, ...);
PetscErrorCode PetscErrorPrintf(const char* fmt, ...);
PetscErrorCode PetscOptionsGetInt(void* opts, const char* pre, const char* name, PetscInt* v, PetscBool* set);
PetscErrorCode PetscOptionsGetReal(void* opts, const char* pre, const char* name, PetscReal* v, PetscBool* set);
PetscErrorCode PetscOptionsGetBool(void* opts, const char* pre, const char* name, PetscBool* v, PetscBool* set);
PetscErrorCode PetscOptionsGetString(void* opts, const char* pre, const char* name, char* v, size_t len, PetscBool* set);
PetscErrorCode PetscObjectTypeCompare(PetscObject obj, const char* type, PetscBool* same);

/* Vec */
PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec* v);
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
PetscErrorCode VecSetValueLocal(Vec v, PetscInt i, PetscScalar a, InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
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
PetscErrorCode VecSetRandom(Vec v, PetscRandom r);
PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom* r);
PetscErrorCode PetscRandomSetType(PetscRandom r, PetscRandomType t);
PetscErrorCode PetscRandomDestroy(PetscRandom* r);

/* viewers */
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm comm, const char* name, PetscFileMode mode, PetscViewer* v);
PetscErrorCode PetscViewerCreate(MPI_Comm comm, PetscViewer* v);
PetscErrorCode PetscViewerSetType(PetscViewer v, PetscViewerType t);
PetscErrorCode PetscViewerFileSetMode(PetscViewer v, PetscFileMode mode);
PetscErrorCode PetscViewerFileSetName(PetscViewer v, const char* name);
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char* fmt, ...);
PetscErrorCode PetscViewerDestroy(PetscViewer* v);

/* Mat (AIJ, local numbering == global numbering on one process) */
PetscErrorCode MatCreateAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscInt dnz, const PetscInt* dnnz,
                            PetscInt onz, const PetscInt* onnz, Mat* A);
PetscErrorCode MatSetLocalToGlobalMapping(Mat A, ISLocalToGlobalMapping r, ISLocalToGlobalMapping c);
PetscErrorCode MatSetValuesLocal(Mat A, PetscInt m, const PetscInt* rows, PetscInt n, const PetscInt* cols, const PetscScalar* v,
                                 InsertMode mode);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);
PetscErrorCode MatZeroEntries(Mat A);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r);
PetscErrorCode MatDiagonalSet(Mat A, Vec d, InsertMode mode);
PetscErrorCode MatDestroy(Mat* A);

/* DMDA */
PetscErrorCode DMDACreate3d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz, DMDAStencilType st, PetscInt M,
                            PetscInt N, PetscInt P, PetscInt m, PetscInt n, PetscInt p, PetscInt dof, PetscInt s, const PetscInt* lx,
                            const PetscInt* ly, const PetscInt* lz, DM* da);
PetscErrorCode DMSetFromOptions(DM da);
PetscErrorCode DMSetUp(DM da);
PetscErrorCode DMDestroy(DM* da);
PetscErrorCode DMDASetUniformCoordinates(DM da, PetscReal xmin, PetscReal xmax, PetscReal ymin, PetscReal ymax, PetscReal zmin,
                                         PetscReal zmax);
PetscErrorCode DMDASetElementType(DM da, DMDAElementType t);
PetscErrorCode DMDAGetInfo(DM da, PetscInt* dim, PetscInt* M, PetscInt* N, PetscInt* P, PetscInt* m, PetscInt* n, PetscInt* p,
                           PetscInt* dof, PetscInt* s, DMBoundaryType* bx, DMBoundaryType* by, DMBoundaryType* bz, DMDAStencilType* st);
PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo* info);
PetscErrorCode DMDAGetCorners(DM da, PetscInt* x, PetscInt* y, PetscInt* z, PetscInt* m, PetscInt* n, PetscInt* p);
PetscErrorCode DMDAGetGhostCorners(DM da, PetscInt* x, PetscInt* y, PetscInt* z, PetscInt* m, PetscInt* n, PetscInt* p);
PetscErrorCode DMDAGetOwnershipRanges(DM da, const PetscInt** lx, const PetscInt** ly, const PetscInt** lz);
PetscErrorCode DMDAGetElements(DM da, PetscInt* nel, PetscInt* nen, const PetscInt** e);
PetscErrorCode DMDARestoreElements(DM da, PetscInt* nel, PetscInt* nen, const PetscInt** e);
PetscErrorCode DMGetCoordinatesLocal(DM da, Vec* c);
PetscErrorCode DMCreateGlobalVector(DM da, Vec* v);
PetscErrorCode DMCreateLocalVector(DM da, Vec* v);
PetscErrorCode DMGlobalToLocalBegin(DM da, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMGlobalToLocalEnd(DM da, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMCreateMatrix(DM da, Mat* A);
PetscErrorCode DMGetLocalToGlobalMapping(DM da, ISLocalToGlobalMapping* m);
PetscErrorCode DMCoarsenHierarchy(DM da, PetscInt nlevels, DM* dac);
PetscErrorCode DMCreateInterpolation(DM dac, DM daf, Mat* R, Vec* scale);

/* KSP / PC */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP* ksp);
PetscErrorCode KSPDestroy(KSP* ksp);
PetscErrorCode KSPSetType(KSP ksp, KSPType t);
PetscErrorCode KSPGetType(KSP ksp, KSPType* t);
PetscErrorCode KSPGMRESSetRestart(KSP ksp, PetscInt r);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal atol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPGetTolerances(KSP ksp, PetscReal* rtol, PetscReal* atol, PetscReal* dtol, PetscInt* maxits);
PetscErrorCode KSPSetInitialGuessNonzero(KSP ksp, PetscBool flg);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P);
PetscErrorCode KSPSetFromOptions(KSP ksp);
PetscErrorCode KSPSetUp(KSP ksp);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt* its);
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal* rnorm);
PetscErrorCode KSPGetPC(KSP ksp, PC* pc);
PetscErrorCode PCSetType(PC pc, PCType t);
PetscErrorCode PCGetType(PC pc, PCType* t);
PetscErrorCode PCSetReusePreconditioner(PC pc, PetscBool flg);
PetscErrorCode PCMGSetLevels(PC pc, PetscInt levels, MPI_Comm* comms);
PetscErrorCode PCMGSetType(PC pc, PCMGType t);
PetscErrorCode PCMGSetCycleType(PC pc, PCMGCycleType t);
PetscErrorCode PCMGSetGalerkin(PC pc, PCMGGalerkinType t);
PetscErrorCode PCMGSetInterpolation(PC pc, PetscInt l, Mat R);
PetscErrorCode PCMGGetSmoother(PC pc, PetscInt l, KSP* ksp);
PetscErrorCode PCMGGetCoarseSolve(PC pc, KSP* ksp);
#endif
