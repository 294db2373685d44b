/* topopt_b200.h -- C-ABI of libtopopt_b200.so
 *
 * B200-native (sm_100a) replacement for the inside of the reference's
 * LinearElasticity / Filter / PDEFilt classes (and, as the "next" row, MMA).
 * Every entry point cites the reference interface (file:line, paths relative to
 * the reference checkout topopt/TopOpt_in_PETSc) it replaces.  No C++ / torch
 * types cross this boundary: plain pointers, sizes and opaque handles.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error (the reference's
 *    PetscErrorCode convention, main.cc:49,65,78); tob_last_error() gives text.
 *  - all ranks call every function collectively, in the same order (SPMD, as
 *    the reference does over PETSC_COMM_WORLD).
 *  - "vec" pointers may be HOST or DEVICE pointers (detected with
 *    cudaPointerGetAttributes).  Host pointers are staged through device
 *    buffers inside the call (this is the end-to-end path main.cc/MMA.cc use
 *    through VecGetArray); device pointers are used in place.
 *  - element vectors have the local size of this rank's owned elements, in the
 *    natural order i + nex*(j + ney*k_local); nodal dof vectors have
 *    3*owned_nodes entries ordered 3*node + c (LinearElasticity.cc:266,819-826).
 *    With the z-slab partition natural == PETSc global ordering.
 *  - node counts nxn,nyn,nzn are the reference's -nx -ny -nz (TopOpt.cc:106-108).
 */
#ifndef TOPOPT_B200_H
#define TOPOPT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tob_ctx tob_ctx;       /* grid + z-slab partition + stream + NCCL comm (replaces DM da_nodes/da_elem, TopOpt.cc:233-300) */
typedef struct tob_elast tob_elast;   /* replaces class LinearElasticity (LinearElasticity.h:21-109) */
typedef struct tob_filter tob_filter; /* replaces class Filter + PDEFilt (Filter.h:34-92, PDEFilter.h:29-74) */
typedef struct tob_mma tob_mma;       /* device MMA behind the MMA class surface (MMA.h:29-140) */

typedef struct tob_grid_desc {
    int nxn, nyn, nzn;   /* global node counts (-nx -ny -nz) */
    double xc[6];        /* xmin,xmax,ymin,ymax,zmin,zmax (TopOpt.cc:109-114) */
    int rank, nranks;    /* z-slab index / number of slabs (one GPU each) */
    int device;          /* CUDA device ordinal, -1 = rank % device_count */
    const void* nccl_id; /* 128-byte ncclUniqueId shared by all ranks; NULL iff nranks==1 */
    void* stream;        /* cudaStream_t to launch on, NULL = library-owned stream */
} tob_grid_desc;

typedef struct tob_solver_opts {
    int pc;              /* 0 = Jacobi, 1 = geometric multigrid V-cycle (replaces PCMG, LinearElasticity.cc:656-746) */
    int nlvls;           /* multigrid levels (-nlvls, TopOpt.cc:116) */
    double rtol, atol, dtol; /* KSPSetTolerances, LinearElasticity.cc:621-651 (1e-5, 1e-50, 1e5) */
    int maxits;          /* 200 for the reference's FGMRES+MG; Jacobi needs far more */
    int check_every;     /* host convergence check cadence (iterations) */
    int cheb_degree;     /* Chebyshev-Jacobi smoothing steps per sweep (reference: GMRES(4)+SOR, 4 its) */
    int coarse_its;      /* coarse-grid PCG iteration cap (reference: GMRES(30)+SOR, 30 its) */
    int spmv_kernel;     /* 0 = auto (7).  1 = naive dense gather (cross-check).  z-march kernels, 32 x TY node columns per CTA,
                            carried state in registers: 5 = TY 8, 9 = TY 16, 4 = TY 12 with one __syncthreads per z-step;
                            6 / 7 / 8 = TY 8 / 16 / 12 with mbarrier row hand-off and a TMA producer warp.
                            2 / 3 = retired shared-memory-state variants, run as 5 / 9. */
    int verbose;
} tob_solver_opts;

typedef struct tob_solve_stats {
    int iterations;
    double rel_residual; /* ||r||/||b||, the "rerr" printed at LinearElasticity.cc:220 */
    double seconds;      /* device time of the solve (CUDA events) */
    int spmv_calls;
    int kernel_launches; /* launches of this library's kernels since the last reset */
    double spmv_seconds; /* summed device time of the fine-level SpMV launches timed with CUDA events (profiling on) */
    int spmv_timed;      /* number of SpMV launches in spmv_seconds */
} tob_solve_stats;

const char* tob_last_error(void);
int tob_version(void);

/* ---- context / grid -------------------------------------------------- */
int tob_ctx_create(const tob_grid_desc* desc, tob_ctx** out);
int tob_ctx_destroy(tob_ctx* ctx);
/* counts: nel = (nx-1)(ny-1)(nz-1), ndof = 3 nx ny nz (TopOpt.cc:172-175) */
int tob_ctx_sizes(tob_ctx* ctx, long long* nel_owned, long long* ndof_owned, long long* nel_global,
                  long long* ndof_global);
/* owned global z ranges of this slab: node planes [kn0,kn1), element layers [ke0,ke1) (LinearElasticity.cc:804-818) */
int tob_ctx_slab(tob_ctx* ctx, int* kn0, int* kn1, int* ke0, int* ke1);
/* the same ownership rule as a pure host function (no device needed): what rank `rank` of `nranks` z-slabs owns
 * on a grid with nzn node planes.  Replaces the DMDA ownership ranges the reference reads with DMDAGetCorners. */
int tob_slab_partition(int nzn, int nranks, int rank, int* kn0, int* kn1, int* ke0, int* ke1);
int tob_ctx_sync(tob_ctx* ctx);
long long tob_ctx_launch_count(tob_ctx* ctx, int reset);
void* tob_ctx_stream(tob_ctx* ctx);
/* device scratch helpers so callers (tests, bench, the C++ Vec shim) can keep vectors resident */
int tob_dev_alloc(tob_ctx* ctx, long long nbytes, void** out);
int tob_dev_free(tob_ctx* ctx, void* p);
int tob_memcpy(tob_ctx* ctx, void* dst, const void* src, long long nbytes); /* any direction, on the ctx stream, synchronous */
/* 128-byte ncclUniqueId for tob_grid_desc.nccl_id (rank 0 creates it, the launcher broadcasts it; stands in for MPI_Init) */
int tob_nccl_unique_id(void* out128);
/* VecScale (main.cc:73) on a host or device vector of n doubles */
int tob_vec_scale(tob_ctx* ctx, double* v, double a, long long n);

/* ---- LinearElasticity ------------------------------------------------ */
void tob_solver_opts_default(tob_solver_opts* o);
/* LinearElasticity::LinearElasticity + SetUpLoadAndBC (LinearElasticity.cc:12-180): KE for (dx,dy,dz,nu), cantilever N/RHS */
int tob_elast_create(tob_ctx* ctx, double nu, const tob_solver_opts* opts, tob_elast** out);
int tob_elast_destroy(tob_elast* h);
/* generalised BC/load: fixed[ndof_owned] (1 = clamped) and rhs[ndof_owned] replace the hard-coded cantilever (LinearElasticity.cc:145-171) */
int tob_elast_set_bc_load(tob_elast* h, const unsigned char* fixed, const double* rhs);
/* 576 doubles, row-major ke[j+24*i] (LinearElasticity.cc:841-998) */
int tob_elast_get_ke(tob_elast* h, double* ke576);
/* y = N.(K(xPhys)(N.u)) + (1-N).u  -- the operator assembled at LinearElasticity.cc:487-549, applied matrix-free */
int tob_elast_set_density(tob_elast* h, const double* xPhys, double Emin, double Emax, double penal);
int tob_elast_spmv(tob_elast* h, const double* u, double* y);
int tob_elast_diag(tob_elast* h, double* d);
/* repeated application on resident vectors for benchmarking: returns mean ms per application (CUDA events) */
int tob_elast_spmv_bench(tob_elast* h, int kernel, int reps, double* ms_per_apply);
/* SolveState (LinearElasticity.cc:182-223): warm-started solve of K U = N.RHS; U stays resident */
int tob_elast_solve(tob_elast* h, const double* xPhys, double Emin, double Emax, double penal);
/* ComputeObjectiveConstraintsSensitivities (LinearElasticity.cc:363-445) */
int tob_elast_solve_obj_sens(tob_elast* h, const double* xPhys, double Emin, double Emax, double penal, double volfrac,
                             double* fx, double* gx, double* dfdx, double* dgdx);
/* ComputeObjectiveConstraints (:225-297) / ComputeSensitivities (:299-361) */
int tob_elast_obj(tob_elast* h, const double* xPhys, double Emin, double Emax, double penal, double volfrac,
                  double* fx, double* gx);
int tob_elast_sens(tob_elast* h, const double* xPhys, double Emin, double Emax, double penal, double volfrac,
                   double* dfdx, double* dgdx);
/* GetStateField (LinearElasticity.h:49): copy U (owned dofs, natural order) out / in (restart, LinearElasticity.cc:555-612) */
int tob_elast_get_state(tob_elast* h, double* U);
int tob_elast_set_state(tob_elast* h, const double* U);
int tob_elast_get_rhs(tob_elast* h, double* rhs);
int tob_elast_last_stats(tob_elast* h, tob_solve_stats* s);
int tob_elast_set_opts(tob_elast* h, const tob_solver_opts* opts);
/* on = 1: time every fine-level SpMV launch of the next solves with CUDA event pairs (tob_solve_stats.spmv_seconds,
 * spmv_timed); on = k > 1: time every k-th launch; 0 = off */
int tob_elast_set_profile(tob_elast* h, int on);

/* ---- Filter / PDEFilt ------------------------------------------------ */
/* Filter::Filter + SetUp (Filter.cc:25-38,290-463); type 0 sens, 1 dens, 2 PDE, else none; rmin physical length */
int tob_filter_create(tob_ctx* ctx, int filter_type, double rmin, tob_filter** out);
int tob_filter_destroy(tob_filter* h);
int tob_filter_elemconn(tob_filter* h, int* elemconn);
/* Filter::FilterProject (Filter.cc:60-117) */
int tob_filter_project(tob_filter* h, const double* x, double* xTilde, double* xPhys, int projection, double beta,
                       double eta);
/* Filter::Gradients (Filter.cc:120-204); dgdx = array of m pointers */
int tob_filter_gradients(tob_filter* h, const double* x, const double* xTilde, double* dfdx, int m, double** dgdx,
                         int projection, double beta, double eta);
/* Filter::GetMND (Filter.cc:206-225) */
int tob_filter_mnd(tob_filter* h, const double* xPhys, double* mnd);
/* Filter::IncreaseBeta (Filter.cc:268-288): pure host logic, returns changeBeta in *changed */
int tob_filter_increase_beta(double* beta, double beta_final, double gx, int itr, double ch, int* changed);
/* Hs = H 1 (Filter.cc:445-448) for tests */
int tob_filter_get_hs(tob_filter* h, double* hs);
/* PDEFilt::FilterProject (PDEFilter.cc:189-216), exposed for tests (type 2 only) */
int tob_filter_pde_apply(tob_filter* h, const double* x, double* out);
int tob_filter_last_stats(tob_filter* h, tob_solve_stats* s);

/* ---- MMA (next row 8f-1) --------------------------------------------- */
/* MMA::MMA(n,m,x[,a,c,d]) (MMA.cc:190-291); n = global design variables, nloc taken from ctx */
int tob_mma_create(tob_ctx* ctx, int m, const double* a, const double* c, const double* d, tob_mma** out);
int tob_mma_destroy(tob_mma* h);
/* MMA::SetOuterMovelimit (MMA.cc:386-405) */
int tob_mma_set_outer_movelimit(tob_mma* h, double Xmin, double Xmax, double movlim, const double* x, double* xmin,
                                double* xmax);
/* MMA::Update (MMA.cc:499-518): x updated in place */
int tob_mma_update(tob_mma* h, double* x, const double* dfdx, const double* gx, double** dgdx, const double* xmin,
                   const double* xmax);
/* MMA::DesignChange (MMA.cc:407-426): inf-norm of x-xold, then xold <- x */
int tob_mma_design_change(tob_mma* h, const double* x, double* xold, double* ch);
/* MMA::Restart (MMA.cc:332-371) / restart ctor (MMA.cc:22-106) */
int tob_mma_get_restart(tob_mma* h, double* xo1, double* xo2, double* U, double* L, int* k);
int tob_mma_set_restart(tob_mma* h, const double* xo1, const double* xo2, const double* U, const double* L, int k);

/* Single-precision variant of tob_elast_spmv on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulation
 * in tensor memory): the smoother-grade operator ("mixed-precision smoother", BASELINE configs[4]).  Same operator
 * (LinearElasticity.cc:510-542 + MatMult), TF32 input rounding: relative accuracy ~1e-3.  u / y fp64 at the boundary. */
int tob_elast_spmv_tf32(tob_elast* h, const double* u, double* y);
int tob_elast_spmv_tf32_bench(tob_elast* h, int reps, double* ms_per_apply);

/* ---- diagnostics for the parity tests -------------------------------------------------------------------------------- */
/* Levels of the multigrid hierarchy (replaces PCMGSetLevels, LinearElasticity.cc:691) and, for a Galerkin level (the R A P
 * product of PCMGSetGalerkin, :702), its assembled 27-point block stencil A[243][nz*ny*px], k = 9*offset + 3*c + c2;
 * dims = {nx, ny, nz, px}.  out may be NULL to query dims / is_stencil only. */
int tob_elast_gmg_nlevels(tob_elast* h, int* nlevels);
int tob_elast_gmg_level_stencil(tob_elast* h, int level, double* out, int* dims, int* is_stencil);

/* ---- micro-benchmarks used to size the kernels (profiles/) ------------ */
/* kind 0: DFMA stream, 1: DMMA m8n8k4 stream, 2: both interleaved; returns TFLOP/s (fp64) */
int tob_microbench_fp64(tob_ctx* ctx, int kind, double* tflops);
/* device-to-device copy bandwidth GB/s (read+write) over nbytes */
int tob_microbench_copy(tob_ctx* ctx, long long nbytes, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* TOPOPT_B200_H */
