// TopOpt.cc -- options, banners, grid handle and design vectors behind the reference's TopOpt surface.
// Behaviour follows TopOpt.cc:102-570 of the reference (defaults, flag names, multigrid divisibility check,
// restart file formats); the mesh itself is a tob_ctx, not a pair of DMDAs.
#include "TopOpt.h"
#include <cmath>
#include <vector>
#include "../../include/topopt_b200.h"

TopOpt::TopOpt(PetscInt nconstraints) {
    m = nconstraints;
    Defaults();
    BuildMesh();
    BuildDesignVectors();
}
TopOpt::TopOpt() : TopOpt(1) {}

void TopOpt::Defaults() {
    x = xTilde = xPhys = dfdx = xmin = xmax = xold = NULL;
    dgdx = NULL; gx = NULL; da_nodes = da_elem = NULL;
    xo1 = xo2 = U = L = NULL;
    // TopOpt.cc:105-135
    nxyz[0] = 65; nxyz[1] = 33; nxyz[2] = 33;
    xc[0] = 0.0; xc[1] = 2.0; xc[2] = 0.0; xc[3] = 1.0; xc[4] = 0.0; xc[5] = 1.0;
    nu = 0.3; nlvls = 4;
    volfrac = 0.12; maxItr = 400; rmin = 0.08; penal = 3.0; Emin = 1.0e-9; Emax = 1.0;
    filter = 1; Xmin = 0.0; Xmax = 1.0; movlim = 0.2;
    restart = PETSC_TRUE; flip = PETSC_TRUE;
    projectionFilter = PETSC_FALSE; beta = 0.1; betaFinal = 48; eta = 0.0;
    fx = 0; fscale = 1; n = nloc = 0;
}

TopOpt::~TopOpt() {
    Vec* vs[] = {&x, &xTilde, &xPhys, &dfdx, &xmin, &xmax, &xold, &xo1, &xo2, &U, &L};
    for (Vec* v : vs) if (*v) VecDestroy(v);
    if (dgdx) VecDestroyVecs(m, &dgdx);
    delete[] gx;
    if (da_elem) DMDestroy(&da_elem);
    if (da_nodes) DMDestroy(&da_nodes);
}

PetscErrorCode TopOpt::BuildMesh() {
    PetscBool flg;
    const char* names[3] = {"-nx", "-ny", "-nz"};
    for (int d = 0; d < 3; d++) PetscOptionsGetInt(NULL, NULL, names[d], &nxyz[d], &flg);
    const char* cn[6] = {"-xcmin", "-xcmax", "-ycmin", "-ycmax", "-zcmin", "-zcmax"};
    for (int d = 0; d < 6; d++) PetscOptionsGetReal(NULL, NULL, cn[d], &xc[d], &flg);
    PetscOptionsGetReal(NULL, NULL, "-penal", &penal, &flg);
    PetscOptionsGetInt(NULL, NULL, "-nlvls", &nlvls, &flg);

    PetscPrintf(PETSC_COMM_WORLD, "########################################################################\n");
    PetscPrintf(PETSC_COMM_WORLD, "############################ FEM settings ##############################\n");
    PetscPrintf(PETSC_COMM_WORLD, "# Number of nodes: (-nx,-ny,-nz):        (%i,%i,%i) \n", nxyz[0], nxyz[1], nxyz[2]);
    PetscPrintf(PETSC_COMM_WORLD, "# Number of degree of freedom:           %i \n", 3 * nxyz[0] * nxyz[1] * nxyz[2]);
    PetscPrintf(PETSC_COMM_WORLD, "# Number of elements:                    (%i,%i,%i) \n", nxyz[0] - 1, nxyz[1] - 1, nxyz[2] - 1);
    PetscPrintf(PETSC_COMM_WORLD, "# Dimensions: (-xcmin,-xcmax,..,-zcmax): (%f,%f,%f)\n", xc[1] - xc[0], xc[3] - xc[2], xc[5] - xc[4]);
    PetscPrintf(PETSC_COMM_WORLD, "# -nlvls: %i\n", nlvls);
    PetscPrintf(PETSC_COMM_WORLD, "########################################################################\n");

    // the element counts must be divisible by 2^(nlvls-1) (TopOpt.cc:182-201)
    const int div = 1 << (nlvls - 1);
    const char* axis[3] = {"X", "Y", "Z"};
    for (int d = 0; d < 3; d++)
        if ((nxyz[d] - 1) % div != 0) {
            PetscPrintf(PETSC_COMM_WORLD, "MESH DIMENSION NOT COMPATIBLE WITH NUMBER OF MULTIGRID LEVELS!\n");
            PetscPrintf(PETSC_COMM_WORLD, "%s - number of nodes %i is cannot be halfened %i times\n", axis[d], nxyz[d], nlvls - 1);
            exit(0);
        }
    dx = (xc[1] - xc[0]) / (PetscScalar)(nxyz[0] - 1);
    dy = (xc[3] - xc[2]) / (PetscScalar)(nxyz[1] - 1);
    dz = (xc[5] - xc[4]) / (PetscScalar)(nxyz[2] - 1);
    PetscErrorCode ierr = DMCreateTob(nxyz[0], nxyz[1], nxyz[2], xc, 1, &da_nodes);
    CHKERRQ(ierr);
    return DMCloneTob(da_nodes, 1, &da_elem);
}

PetscErrorCode TopOpt::BuildDesignVectors() {
    long long nel_own = 0, nel_glob = 0;
    tob_ctx_sizes(DMGetTobCtx(da_nodes), &nel_own, NULL, &nel_glob, NULL);
    nloc = (PetscInt)nel_own;
    VecCreateTob(nloc, &xPhys);
    VecGetSize(xPhys, &n);
    PetscBool flg;
    PetscOptionsGetReal(NULL, NULL, "-Emin", &Emin, &flg);
    PetscOptionsGetReal(NULL, NULL, "-Emax", &Emax, &flg);
    PetscOptionsGetReal(NULL, NULL, "-nu", &nu, &flg);
    PetscOptionsGetReal(NULL, NULL, "-volfrac", &volfrac, &flg);
    PetscOptionsGetReal(NULL, NULL, "-penal", &penal, &flg);
    PetscOptionsGetReal(NULL, NULL, "-rmin", &rmin, &flg);
    PetscOptionsGetInt(NULL, NULL, "-maxItr", &maxItr, &flg);
    PetscOptionsGetInt(NULL, NULL, "-filter", &filter, &flg);
    PetscOptionsGetReal(NULL, NULL, "-Xmin", &Xmin, &flg);
    PetscOptionsGetReal(NULL, NULL, "-Xmax", &Xmax, &flg);
    PetscOptionsGetReal(NULL, NULL, "-movlim", &movlim, &flg);
    PetscOptionsGetBool(NULL, NULL, "-projectionFilter", &projectionFilter, &flg);
    PetscOptionsGetReal(NULL, NULL, "-beta", &beta, &flg);
    PetscOptionsGetReal(NULL, NULL, "-betaFinal", &betaFinal, &flg);
    PetscOptionsGetReal(NULL, NULL, "-eta", &eta, &flg);

    PetscPrintf(PETSC_COMM_WORLD, "################### Optimization settings ####################\n");
    PetscPrintf(PETSC_COMM_WORLD, "# Problem size: n= %i, m= %i\n", n, m);
    PetscPrintf(PETSC_COMM_WORLD, "# -filter: %i  (0=sens., 1=dens, 2=PDE)\n", filter);
    PetscPrintf(PETSC_COMM_WORLD, "# -rmin: %f\n", rmin);
    PetscPrintf(PETSC_COMM_WORLD, "# -projectionFilter: %i  (0/1)\n", projectionFilter);
    PetscPrintf(PETSC_COMM_WORLD, "# -beta: %f\n", beta);
    PetscPrintf(PETSC_COMM_WORLD, "# -betaFinal: %f\n", betaFinal);
    PetscPrintf(PETSC_COMM_WORLD, "# -eta: %f\n", eta);
    PetscPrintf(PETSC_COMM_WORLD, "# -volfrac: %f\n", volfrac);
    PetscPrintf(PETSC_COMM_WORLD, "# -penal: %f\n", penal);
    PetscPrintf(PETSC_COMM_WORLD, "# -Emin/-Emax: %e - %e \n", Emin, Emax);
    PetscPrintf(PETSC_COMM_WORLD, "# -nu: %f \n", nu);
    PetscPrintf(PETSC_COMM_WORLD, "# -maxItr: %i\n", maxItr);
    PetscPrintf(PETSC_COMM_WORLD, "# -movlim: %f\n", movlim);
    PetscPrintf(PETSC_COMM_WORLD, "##############################################################\n");

    gx = new PetscScalar[m];
    if (filter == 0) Xmin = 0.001;  // TopOpt.cc:357-359
    Vec* mk[] = {&x, &xTilde, &dfdx, &xmin, &xmax, &xold};
    for (Vec* v : mk) VecDuplicate(xPhys, v);
    VecDuplicateVecs(xPhys, m, &dgdx);
    VecSet(x, volfrac); VecSet(xTilde, volfrac); VecSet(xPhys, volfrac); VecSet(xold, volfrac);
    return 0;
}

PetscErrorCode TopOpt::AllocateMMAwithRestart(PetscInt* itr, MMA** mma) {
    std::vector<PetscScalar> a(m, 0.0), c(m, 1000.0), d(m, 0.0);  // TopOpt.cc:394-398
    restart = PETSC_TRUE; flip = PETSC_TRUE;
    PetscBool onlyLoadDesign = PETSC_FALSE, flg;
    char buf[PETSC_MAX_PATH_LEN];
    PetscOptionsGetBool(NULL, NULL, "-restart", &restart, &flg);
    PetscOptionsGetBool(NULL, NULL, "-onlyLoadDesign", &onlyLoadDesign, &flg);
    if (restart) { Vec* vs[] = {&xo1, &xo2, &U, &L}; for (Vec* v : vs) VecDuplicate(x, v); }
    std::string workdir = "./";
    PetscOptionsGetString(NULL, NULL, "-workdir", buf, sizeof(buf), &flg);
    if (flg) workdir = buf;
    file00 = workdir + "/Restart00.dat"; file00Itr = workdir + "/Restart00_itr_f0.dat";
    file01 = workdir + "/Restart01.dat"; file01Itr = workdir + "/Restart01_itr_f0.dat";
    std::string fvec, fitr;
    PetscOptionsGetString(NULL, NULL, "-restartFileVec", buf, sizeof(buf), &flg);
    if (flg) fvec = buf;
    PetscOptionsGetString(NULL, NULL, "-restartFileItr", buf, sizeof(buf), &flg);
    if (flg) fitr = buf;
    PetscPrintf(PETSC_COMM_WORLD, "##############################################################\n");
    PetscPrintf(PETSC_COMM_WORLD, "# Continue from previous iteration (-restart): %i \n", restart);
    PetscPrintf(PETSC_COMM_WORLD, "# Restart file (-restartFileVec): %s \n", fvec.c_str());
    PetscPrintf(PETSC_COMM_WORLD, "# Restart file (-restartFileItr): %s \n", fitr.c_str());
    PetscPrintf(PETSC_COMM_WORLD, "# New restart files are written to (-workdir): %s (Restart0x.dat and Restart0x_itr_f0.dat) \n", workdir.c_str());
    const bool haveVec = Exists(fvec), haveItr = Exists(fitr);
    if (!haveVec) PetscPrintf(PETSC_COMM_WORLD, "File: %s NOT FOUND \n", fvec.c_str());
    if (!haveItr) PetscPrintf(PETSC_COMM_WORLD, "File: %s NOT FOUND \n", fitr.c_str());
    PetscInt nglob;
    VecGetSize(x, &nglob);
    if (restart && haveVec && haveItr) {
        PetscViewer view;
        PetscErrorCode ierr = PetscViewerBinaryOpen(PETSC_COMM_WORLD, fvec.c_str(), FILE_MODE_READ, &view);
        CHKERRQ(ierr);
        Vec order[] = {x, xPhys, xo1, xo2, U, L};  // stream order of TopOpt.cc:481-486
        for (Vec v : order) { ierr = VecLoad(v, view); CHKERRQ(ierr); }
        PetscViewerDestroy(&view);
        std::fstream f(fitr.c_str(), std::ios_base::in);
        f >> itr[0];
        f >> fscale;
        if (onlyLoadDesign) {
            PetscPrintf(PETSC_COMM_WORLD, "# Loading design from file: %s \n", fvec.c_str());
            *mma = new MMA(nglob, m, x, a.data(), c.data(), d.data());
        } else {
            PetscPrintf(PETSC_COMM_WORLD, "# Continue optimization from file: %s \n", fvec.c_str());
            *mma = new MMA(nglob, m, *itr, xo1, xo2, U, L, a.data(), c.data(), d.data());
        }
        PetscPrintf(PETSC_COMM_WORLD, "# Successful restart from file: %s and %s \n", fvec.c_str(), fitr.c_str());
    } else {
        *mma = new MMA(nglob, m, x, a.data(), c.data(), d.data());
    }
    return 0;
}

PetscErrorCode TopOpt::WriteRestartFiles(PetscInt* itr, MMA* mma) {
    if (!restart) return -1;
    mma->Restart(xo1, xo2, U, L);
    flip = flip ? PETSC_FALSE : PETSC_TRUE;  // alternate between the two dump streams
    const std::string& fv = flip ? file01 : file00;
    const std::string& fi = flip ? file01Itr : file00Itr;
    PetscViewer view, txt;
    PetscViewerBinaryOpen(PETSC_COMM_WORLD, fv.c_str(), FILE_MODE_WRITE, &view);
    PetscViewerASCIIOpen(PETSC_COMM_WORLD, fi.c_str(), &txt);
    PetscViewerASCIIPrintf(txt, "%d ", itr[0]);
    PetscViewerASCIIPrintf(txt, " %e", fscale);  // 7 significant digits, as the reference (TopOpt.cc:553)
    PetscViewerASCIIPrintf(txt, "\n");
    Vec order[] = {x, xPhys, xo1, xo2, U, L};  // TopOpt.cc:556-562
    for (Vec v : order) VecView(v, view);
    PetscViewerDestroy(&view);
    PetscViewerDestroy(&txt);
    return 0;
}
