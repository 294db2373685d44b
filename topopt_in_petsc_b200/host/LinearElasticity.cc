// LinearElasticity.cc -- forwarding wrapper: every Compute* call hands the caller's host arrays
// (VecGetArray) to libtopopt_b200.so, which stages them to the GPU, solves and returns the results.
#include "LinearElasticity.h"
#include <fstream>
#include "../../include/topopt_b200.h"

LinearElasticity::LinearElasticity(DM da_nodes) : da_nodal(NULL), h(NULL), U(NULL), state_loaded(false) {
    nu = 0.3; nlvls = 4;  // LinearElasticity.cc:22-26
    PetscBool flg;
    PetscOptionsGetInt(NULL, NULL, "-nlvls", &nlvls, &flg);
    PetscOptionsGetReal(NULL, NULL, "-nu", &nu, &flg);
    tob_solver_opts o;
    tob_solver_opts_default(&o);
    o.nlvls = nlvls;
    // solver selection flags of this build (the reference's -ksp_* / -pc_* have no meaning here)
    PetscOptionsGetInt(NULL, NULL, "-tob_pc", &o.pc, &flg);
    PetscOptionsGetReal(NULL, NULL, "-tob_rtol", &o.rtol, &flg);
    PetscOptionsGetInt(NULL, NULL, "-tob_maxits", &o.maxits, &flg);
    PetscOptionsGetInt(NULL, NULL, "-tob_cheb", &o.cheb_degree, &flg);
    PetscOptionsGetInt(NULL, NULL, "-tob_spmv", &o.spmv_kernel, &flg);
    DMCloneTob(da_nodes, 3, &da_nodal);
    if (tob_elast_create(DMGetTobCtx(da_nodal), nu, &o, &h) != 0) {
        PetscErrorPrintf("tob_elast_create: %s\n", tob_last_error());
        MPI_Abort(PETSC_COMM_WORLD, 77);
    }
    long long ndof_own = 0;
    tob_ctx_sizes(DMGetTobCtx(da_nodal), NULL, &ndof_own, NULL, NULL);
    VecCreateTob((PetscInt)ndof_own, &U);
    restart = PETSC_TRUE; flip = PETSC_TRUE;
    PetscOptionsGetBool(NULL, NULL, "-restart", &restart, &flg);
    std::string workdir = "./";
    char buf[PETSC_MAX_PATH_LEN];
    PetscOptionsGetString(NULL, NULL, "-workdir", buf, sizeof(buf), &flg);
    if (flg) workdir = buf;
    file00 = workdir + "/RestartSol00.dat";
    file01 = workdir + "/RestartSol01.dat";
}

LinearElasticity::~LinearElasticity() {
    if (h) tob_elast_destroy(h);
    VecDestroy(&U);
    if (da_nodal) DMDestroy(&da_nodal);
}

void LinearElasticity::LoadRestartState() {
    // the reference reads the state restart lazily inside SetUpSolver (LinearElasticity.cc:555-612)
    state_loaded = true;
    PetscBool flg, onlyDesign = PETSC_FALSE;
    PetscOptionsGetBool(NULL, NULL, "-onlyLoadDesign", &onlyDesign, &flg);
    if (!restart || onlyDesign) return;
    char buf[PETSC_MAX_PATH_LEN];
    std::string f;
    PetscOptionsGetString(NULL, NULL, "-restartFileVecSol", buf, sizeof(buf), &flg);
    if (flg) f = buf;
    PetscPrintf(PETSC_COMM_WORLD, "# Restarting with solution (State Vector) from (-restartFileVecSol): %s \n", f.c_str());
    std::ifstream test(f.c_str());
    if (!test) { PetscPrintf(PETSC_COMM_WORLD, "File: %s NOT FOUND \n", f.c_str()); return; }
    PetscViewer view;
    if (PetscViewerBinaryOpen(PETSC_COMM_WORLD, f.c_str(), FILE_MODE_READ, &view) == 0 && VecLoad(U, view) == 0) {
        PetscScalar* up;
        VecGetArray(U, &up);
        tob_elast_set_state(h, up);
    }
    PetscViewerDestroy(&view);
}

PetscErrorCode LinearElasticity::Report(int rc, bool solved) {
    if (rc != 0) { PetscErrorPrintf("libtopopt_b200: %s\n", tob_last_error()); return rc; }
    if (solved) {
        tob_solve_stats s;
        tob_elast_last_stats(h, &s);
        PetscPrintf(PETSC_COMM_WORLD, "State solver:  iter: %i, rerr.: %e, time: %f\n", s.iterations, s.rel_residual, s.seconds);
    }
    return 0;
}

PetscErrorCode LinearElasticity::ComputeObjectiveConstraintsSensitivities(PetscScalar* fx, PetscScalar* gx, Vec dfdx, Vec dgdx,
                                                                          Vec xPhys, PetscScalar Emin, PetscScalar Emax,
                                                                          PetscScalar penal, PetscScalar volfrac) {
    if (!state_loaded) LoadRestartState();
    PetscScalar *xp, *df, *dg;
    VecGetArray(xPhys, &xp); VecGetArray(dfdx, &df); VecGetArray(dgdx, &dg);
    return Report(tob_elast_solve_obj_sens(h, xp, Emin, Emax, penal, volfrac, fx, gx, df, dg), true);
}
PetscErrorCode LinearElasticity::ComputeObjectiveConstraints(PetscScalar* fx, PetscScalar* gx, Vec xPhys, PetscScalar Emin,
                                                             PetscScalar Emax, PetscScalar penal, PetscScalar volfrac) {
    if (!state_loaded) LoadRestartState();
    PetscScalar* xp;
    VecGetArray(xPhys, &xp);
    return Report(tob_elast_obj(h, xp, Emin, Emax, penal, volfrac, fx, gx), true);
}
PetscErrorCode LinearElasticity::ComputeSensitivities(Vec dfdx, Vec dgdx, Vec xPhys, PetscScalar Emin, PetscScalar Emax,
                                                      PetscScalar penal, PetscScalar volfrac) {
    PetscScalar *xp, *df, *dg;
    VecGetArray(xPhys, &xp); VecGetArray(dfdx, &df); VecGetArray(dgdx, &dg);
    return Report(tob_elast_sens(h, xp, Emin, Emax, penal, volfrac, df, dg), false);
}

Vec LinearElasticity::GetStateField() {
    PetscScalar* up;
    VecGetArray(U, &up);
    if (tob_elast_get_state(h, up) != 0) PetscErrorPrintf("libtopopt_b200: %s\n", tob_last_error());
    return U;
}

PetscErrorCode LinearElasticity::WriteRestartFiles() {
    if (!restart) return -1;  // LinearElasticity.cc:451-453
    flip = flip ? PETSC_FALSE : PETSC_TRUE;
    PetscViewer view;
    PetscViewerBinaryOpen(PETSC_COMM_WORLD, (flip ? file01 : file00).c_str(), FILE_MODE_WRITE, &view);
    VecView(GetStateField(), view);
    PetscViewerDestroy(&view);
    return 0;
}
