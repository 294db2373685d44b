// topopt_b200_main.cc -- the product's own driver: the design loop of the reference (main.cc:48-123)
// over the B200 classes, with the device MMA flavour (host/mma_device) instead of the host MMA.cc.
// The per-iteration line has the reference's format so logs can be diffed against a PETSc run; a summary
// with iterations/sec is printed at the end.  (The drop-in proof -- the reference's untouched main.cc and
// MMA.cc linked against these same classes -- is built by oracle/Makefile into oracle/_ref/topopt_dropin.)
#include <vector>
#include "Filter.h"
#include "LinearElasticity.h"
#include "MMA.h"
#include "MPIIO.h"
#include "TopOpt.h"

struct LoopStats {
    std::vector<double> seconds;
    double total() const { double s = 0; for (double v : seconds) s += v; return s; }
};

static PetscErrorCode design_loop(TopOpt& opt, LinearElasticity& physics, Filter& filter, MPIIO& output, MMA& mma, PetscInt itr,
                                  LoopStats& stats) {
    PetscErrorCode ierr = filter.FilterProject(opt.x, opt.xTilde, opt.xPhys, opt.projectionFilter, opt.beta, opt.eta);
    CHKERRQ(ierr);
    PetscScalar change = 1.0;
    for (; itr < opt.maxItr && change > 0.01;) {
        ++itr;
        const double t0 = MPI_Wtime();
        ierr = physics.ComputeObjectiveConstraintsSensitivities(&opt.fx, &opt.gx[0], opt.dfdx, opt.dgdx[0], opt.xPhys, opt.Emin,
                                                                opt.Emax, opt.penal, opt.volfrac);
        CHKERRQ(ierr);
        if (itr == 1) opt.fscale = 10.0 / opt.fx;  // objective scaling fixed at the first iteration
        opt.fx *= opt.fscale;
        VecScale(opt.dfdx, opt.fscale);
        ierr = filter.Gradients(opt.x, opt.xTilde, opt.dfdx, opt.m, opt.dgdx, opt.projectionFilter, opt.beta, opt.eta);
        CHKERRQ(ierr);
        ierr = mma.SetOuterMovelimit(opt.Xmin, opt.Xmax, opt.movlim, opt.x, opt.xmin, opt.xmax);
        CHKERRQ(ierr);
        ierr = mma.Update(opt.x, opt.dfdx, opt.gx, opt.dgdx, opt.xmin, opt.xmax);
        CHKERRQ(ierr);
        change = mma.DesignChange(opt.x, opt.xold);
        PetscBool betaChanged = PETSC_FALSE;
        if (opt.projectionFilter) betaChanged = filter.IncreaseBeta(&opt.beta, opt.betaFinal, opt.gx[0], itr, change);
        ierr = filter.FilterProject(opt.x, opt.xTilde, opt.xPhys, opt.projectionFilter, opt.beta, opt.eta);
        CHKERRQ(ierr);
        const PetscScalar mnd = filter.GetMND(opt.xPhys);
        const double dt = MPI_Wtime() - t0;
        stats.seconds.push_back(dt);
        PetscPrintf(PETSC_COMM_WORLD, "It.: %i, True fx: %f, Scaled fx: %f, gx[0]: %f, ch.: %f, mnd.: %f, time: %f\n", itr,
                    opt.fx / opt.fscale, opt.fx, opt.gx[0], change, mnd, dt);
        if (itr < 11 || itr % 20 == 0 || betaChanged)
            output.WriteVTK(physics.da_nodal, physics.GetStateField(), opt.x, opt.xTilde, opt.xPhys, itr);
        if (itr % 10 == 0) { opt.WriteRestartFiles(&itr, &mma); physics.WriteRestartFiles(); }
    }
    opt.WriteRestartFiles(&itr, &mma);
    physics.WriteRestartFiles();
    output.WriteVTK(physics.da_nodal, physics.GetStateField(), opt.x, opt.xTilde, opt.xPhys, itr + 1);
    return 0;
}

int main(int argc, char** argv) {
    PetscInitialize(&argc, &argv, PETSC_NULL, "3D topology optimisation, B200-native inner loop");
    LoopStats stats;
    PetscErrorCode ierr = 0;
    {
        TopOpt opt;
        LinearElasticity physics(opt.da_nodes);
        Filter filter(opt.da_nodes, opt.xPhys, opt.filter, opt.rmin);
        MPIIO output(opt.da_nodes, 3, "ux, uy, uz", 3, "x, xTilde, xPhys");
        MMA* mma = NULL;
        PetscInt itr = 0;
        opt.AllocateMMAwithRestart(&itr, &mma);
        ierr = design_loop(opt, physics, filter, output, *mma, itr, stats);
        delete mma;
    }
    if (!stats.seconds.empty())
        PetscPrintf(PETSC_COMM_WORLD, "# %d design iterations in %.3f s  (%.3f iterations/s)\n", (int)stats.seconds.size(),
                    stats.total(), stats.seconds.size() / stats.total());
    PetscFinalize();
    return ierr;
}
