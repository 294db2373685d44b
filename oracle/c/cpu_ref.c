/* cpu_ref.c -- C/OpenMP restatement of the reference's CPU state-solve path (TEST/BASELINE
 * INFRASTRUCTURE ONLY: used by tests/, bench.py's cpu_baseline / --impl reference legs, never by
 * the product path).
 *
 * What it restates (paths relative to the reference checkout):
 *   - Hex8Isoparametric, LinearElasticity.cc:841-998 (2x2x2 Gauss, B^T C B)
 *   - AssembleStiffnessMatrix, LinearElasticity.cc:487-549: K = sum_e E_e KE into an AIJ/CSR
 *     matrix with the DMDA box-stencil pattern (81 nnz per interior row), then
 *     K <- N K N + (I-N)
 *   - the Krylov solve of SolveState (LinearElasticity.cc:182-223) on the assembled matrix, two ways:
 *       cref_pcg    Jacobi-preconditioned CG (the GPU arm's cfg2 algorithm; used by parity tests as a
 *                   second independent solver), and
 *       cref_fgmres_mg  the reference's OWN solver configuration (LinearElasticity.cc:613-746) restated
 *                   from PETSc's documented algorithms (PETSc itself is not installable here):
 *                   FGMRES(100) right-preconditioned, rtol 1e-5 / atol 1e-50 / dtol 1e5 / 200 its, warm
 *                   start; PCMG multiplicative V-cycle on the DMDA hierarchy coarsened by 2, Galerkin
 *                   coarse operators R A P with R = P^T, P = DMDA Q1 (trilinear) interpolation per dof;
 *                   level smoothers GMRES(4) x 4 iterations (no convergence test, as PCMG sets
 *                   KSP_NORM_NONE) left-preconditioned with PCSOR (default: one local symmetric sweep,
 *                   omega = 1, zero initial guess -- "local" = the rows of one process; here one OpenMP
 *                   thread plays one MPI rank, rows split in contiguous natural-order blocks like the
 *                   DMDA's z-slabs); coarse solve GMRES(30) + SOR, rtol 1e-8, <= 30 iterations.
 *       cref_mg_setup   the per-design-iteration KSPSetUp (LinearElasticity.cc:198-200): Galerkin
 *                   products for every coarse level.
 *   - objective / sensitivities, LinearElasticity.cc:403-437.
 *   - the cone density / sensitivity filter as the reference applies it (MatMult with H and the Hs
 *     scaling, Filter.cc:60-117,120-204), H applied from its definition Filter.cc:404-440.
 * PARITY STATUS: pinned through the numpy oracle (tests/test_oracle_cpu.py), which reproduces the output of the
 * reference's own unmodified program run on oracle/petsc_lite (tests/test_reference_cpu.py,
 * tests/golden/reference_runs.json).
 */
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CREF_MAXL 8
/* one multigrid level: box-stencil CSR (81 nnz per interior row) on an nxn x nyn x nzn node grid, 3 dofs per node */
typedef struct {
    int nxn, nyn, nzn;
    long long nnode, ndof, nnz;
    long long* rowptr;
    int* col;
    double* val;
    int* nbr_block;
    double* dinv; /* 1 / diagonal (SOR) */
    double *b, *x, *r, *w, *t; /* V-cycle work vectors */
    double* V;                 /* GMRES basis of the level smoother / coarse solver */
} lvl_t;

typedef struct {
    lvl_t lv[CREF_MAXL]; /* lv[0] = the fine (design) grid; its arrays are also reachable through the fields below */
    int nl;
    int nthreads;
    int nxn, nyn, nzn, nex, ney, nez;
    long long nnode, ndof, nel;
    double dx, dy, dz, nu;
    double KE[576];
    unsigned char* fixed; /* per dof */
    double* rhs;
    /* CSR */
    long long* rowptr;
    int* col;
    double* val;
    long long nnz;
    int* nbr_block; /* per node: index (in units of 3 columns) of each of the 27 neighbour blocks, -1 if absent */
    double* E;
    /* FGMRES(restart) outer Krylov space, allocated as the iteration needs it (PETSc allocates in chunks too) */
    double **fgV, **fgZ;
    int fg_nv, fg_nz;
    long long mg_smooth_calls, mg_coarse_its;
} cref_t;
static void cref_free_krylov(cref_t* h);

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* reference node order (LinearElasticity.cc:118-120) */
static const int SG[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};

void cref_hex8_ke(double dx, double dy, double dz, double nu, double* ke) {
    const double lambda = nu / ((1.0 + nu) * (1.0 - 2.0 * nu)), mu = 1.0 / (2.0 * (1.0 + nu));
    double C[6][6];
    memset(C, 0, sizeof C);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i][j] = lambda + (i == j ? 2.0 * mu : 0.0);
    C[3][3] = C[4][4] = C[5][5] = mu;
    const double GP[2] = {-0.577350269189626, 0.577350269189626};
    memset(ke, 0, 576 * sizeof(double));
    for (int ii = 0; ii < 2; ii++)
        for (int jj = 0; jj < 2; jj++)
            for (int kk = 0; kk < 2; kk++) {
                const double xi = GP[ii], eta = GP[jj], zeta = GP[kk];
                double B[6][24];
                memset(B, 0, sizeof B);
                for (int n = 0; n < 8; n++) {
                    const double gx = 0.125 * SG[n][0] * (1 + SG[n][1] * eta) * (1 + SG[n][2] * zeta) / (0.5 * dx);
                    const double gy = 0.125 * SG[n][1] * (1 + SG[n][0] * xi) * (1 + SG[n][2] * zeta) / (0.5 * dy);
                    const double gz = 0.125 * SG[n][2] * (1 + SG[n][0] * xi) * (1 + SG[n][1] * eta) / (0.5 * dz);
                    B[0][3 * n] = gx; B[1][3 * n + 1] = gy; B[2][3 * n + 2] = gz;
                    B[3][3 * n] = gy; B[3][3 * n + 1] = gx;
                    B[4][3 * n + 1] = gz; B[4][3 * n + 2] = gy;
                    B[5][3 * n] = gz; B[5][3 * n + 2] = gx;
                }
                const double w = dx * dy * dz / 8.0;
                for (int i = 0; i < 24; i++)
                    for (int j = 0; j < 24; j++) {
                        double s = 0;
                        for (int k = 0; k < 6; k++)
                            for (int l = 0; l < 6; l++) s += B[k][i] * C[k][l] * B[l][j];
                        ke[j + 24 * i] += w * s;
                    }
            }
}

static inline long long nid(const cref_t* h, int i, int j, int k) { return i + (long long)h->nxn * (j + (long long)h->nyn * k); }

static inline long long lnid(const lvl_t* L, int i, int j, int k) { return i + (long long)L->nxn * (j + (long long)L->nyn * k); }

/* DMDA box-stencil pattern of one level (DMCreateMatrix on a dof-3 DMDA with DMDA_STENCIL_BOX, width 1) */
static void lvl_pattern(lvl_t* L, int nxn, int nyn, int nzn) {
    memset(L, 0, sizeof *L);
    L->nxn = nxn; L->nyn = nyn; L->nzn = nzn;
    L->nnode = (long long)nxn * nyn * nzn; L->ndof = 3 * L->nnode;
    L->rowptr = (long long*)malloc((L->ndof + 1) * sizeof(long long));
    L->nbr_block = (int*)malloc(L->nnode * 27 * sizeof(int));
    long long* cnt = (long long*)malloc((L->nnode + 1) * sizeof(long long));
    cnt[0] = 0;
    for (int k = 0; k < nzn; k++)
        for (int j = 0; j < nyn; j++)
            for (int i = 0; i < nxn; i++) {
                long long n = lnid(L, i, j, k);
                int c = 0;
                for (int dk = -1; dk <= 1; dk++)
                    for (int dj = -1; dj <= 1; dj++)
                        for (int di = -1; di <= 1; di++) {
                            int ok = i + di >= 0 && i + di < nxn && j + dj >= 0 && j + dj < nyn && k + dk >= 0 && k + dk < nzn;
                            L->nbr_block[n * 27 + (dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] = ok ? c : -1;
                            c += ok;
                        }
                cnt[n + 1] = cnt[n] + c;
            }
    L->nnz = 9 * cnt[L->nnode];
    L->col = (int*)malloc(L->nnz * sizeof(int));
    L->val = (double*)malloc(L->nnz * sizeof(double));
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < L->nnode; n++) {
        const int i = n % nxn, j = (n / nxn) % nyn, k = n / ((long long)nxn * nyn);
        const long long nb = cnt[n + 1] - cnt[n];
        for (int c = 0; c < 3; c++) L->rowptr[3 * n + c] = 9 * cnt[n] + c * 3 * nb;
        for (int dk = -1; dk <= 1; dk++)
            for (int dj = -1; dj <= 1; dj++)
                for (int di = -1; di <= 1; di++) {
                    int b = L->nbr_block[n * 27 + (dk + 1) * 9 + (dj + 1) * 3 + (di + 1)];
                    if (b < 0) continue;
                    long long m = lnid(L, i + di, j + dj, k + dk);
                    for (int c = 0; c < 3; c++)
                        for (int c2 = 0; c2 < 3; c2++) {
                            L->col[L->rowptr[3 * n + c] + 3 * b + c2] = (int)(3 * m + c2);
                            L->val[L->rowptr[3 * n + c] + 3 * b + c2] = 0.0; /* first touch by the owning thread */
                        }
                }
    }
    L->rowptr[L->ndof] = L->nnz;
    free(cnt);
}
static void lvl_free(lvl_t* L) {
    free(L->rowptr); free(L->col); free(L->val); free(L->nbr_block); free(L->dinv);
    free(L->b); free(L->x); free(L->r); free(L->w); free(L->t); free(L->V);
    memset(L, 0, sizeof *L);
}

cref_t* cref_create(int nxn, int nyn, int nzn, double dx, double dy, double dz, double nu, int nthreads) {
    cref_t* h = (cref_t*)calloc(1, sizeof(cref_t));
    if (nthreads > 0) omp_set_num_threads(nthreads);
    h->nthreads = omp_get_max_threads();
    h->nxn = nxn; h->nyn = nyn; h->nzn = nzn; h->nex = nxn - 1; h->ney = nyn - 1; h->nez = nzn - 1;
    h->nnode = (long long)nxn * nyn * nzn; h->ndof = 3 * h->nnode; h->nel = (long long)h->nex * h->ney * h->nez;
    h->dx = dx; h->dy = dy; h->dz = dz; h->nu = nu;
    cref_hex8_ke(dx, dy, dz, nu, h->KE);
    h->fixed = (unsigned char*)calloc(h->ndof, 1);
    h->rhs = (double*)calloc(h->ndof, sizeof(double));
    /* cantilever BC/load, LinearElasticity.cc:143-171 */
    for (int k = 0; k < nzn; k++)
        for (int j = 0; j < nyn; j++) {
            long long n = nid(h, 0, j, k);
            h->fixed[3 * n] = h->fixed[3 * n + 1] = h->fixed[3 * n + 2] = 1;
        }
    for (int j = 0; j < nyn; j++) h->rhs[3 * nid(h, nxn - 1, j, 0) + 2] = (j == 0 || j == nyn - 1) ? -0.0005 : -0.001;
    /* DMDA box-stencil sparsity (DMCreateMatrix, LinearElasticity.cc:105): 27 neighbour nodes x 3 dofs */
    lvl_pattern(&h->lv[0], nxn, nyn, nzn);
    h->nl = 1;
    h->rowptr = h->lv[0].rowptr; h->col = h->lv[0].col; h->val = h->lv[0].val; h->nbr_block = h->lv[0].nbr_block;
    h->nnz = h->lv[0].nnz;
    h->E = (double*)malloc(h->nel * sizeof(double));
    return h;
}


void cref_destroy(cref_t* h) {
    if (!h) return;
    free(h->fixed); free(h->rhs); free(h->E);
    for (int l = 0; l < h->nl; l++) lvl_free(&h->lv[l]);
    cref_free_krylov(h);
    free(h);
}

long long cref_ndof(cref_t* h) { return h->ndof; }
long long cref_nel(cref_t* h) { return h->nel; }
long long cref_nnz(cref_t* h) { return h->nnz; }
void cref_get_ke(cref_t* h, double* ke) { memcpy(ke, h->KE, sizeof h->KE); }
void cref_get_rhs(cref_t* h, double* rhs) { memcpy(rhs, h->rhs, h->ndof * sizeof(double)); }

/* local node of element (ei,ej,ek) sitting at grid node (ei+a, ej+b, ek+g), reference order */
static inline int lnode(int a, int b, int g) { return (g ? 4 : 0) + (b ? (a ? 2 : 3) : (a ? 1 : 0)); }

/* AssembleStiffnessMatrix, LinearElasticity.cc:487-549 (row-wise gather form; same sums). Returns seconds. */
double cref_assemble(cref_t* h, const double* xphys, double Emin, double Emax, double penal) {
    const double t0 = now();
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < h->nel; e++) h->E[e] = Emin + pow(xphys[e], penal) * (Emax - Emin); /* :519 */
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < h->nnode; n++) {
        const int i = n % h->nxn, j = (n / h->nxn) % h->nyn, k = n / ((long long)h->nxn * h->nyn);
        for (int c = 0; c < 3; c++)
            for (long long p = h->rowptr[3 * n + c]; p < h->rowptr[3 * n + c + 1]; p++) h->val[p] = 0.0;
        for (int dk = -1; dk <= 0; dk++)
            for (int dj = -1; dj <= 0; dj++)
                for (int di = -1; di <= 0; di++) {
                    const int ei = i + di, ej = j + dj, ek = k + dk;
                    if (ei < 0 || ei >= h->nex || ej < 0 || ej >= h->ney || ek < 0 || ek >= h->nez) continue;
                    const double Ee = h->E[ei + (long long)h->nex * (ej + (long long)h->ney * ek)];
                    const int ln = lnode(-di, -dj, -dk);
                    for (int g = 0; g < 2; g++)
                        for (int b = 0; b < 2; b++)
                            for (int a = 0; a < 2; a++) {
                                const int lm = lnode(a, b, g);
                                /* neighbour offset of node m relative to n */
                                const int oi = di + a, oj = dj + b, ok = dk + g;
                                const int blk = h->nbr_block[n * 27 + (ok + 1) * 9 + (oj + 1) * 3 + (oi + 1)];
                                for (int c = 0; c < 3; c++) {
                                    double* row = h->val + h->rowptr[3 * n + c] + 3 * blk;
                                    for (int c2 = 0; c2 < 3; c2++) row[c2] += Ee * h->KE[(3 * ln + c) * 24 + 3 * lm + c2];
                                }
                            }
                }
    }
    /* Dirichlet: K <- N K N + (I - N), :532-538 */
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < h->ndof; r++)
        for (long long p = h->rowptr[r]; p < h->rowptr[r + 1]; p++) {
            const int cidx = h->col[p];
            if (h->fixed[r] || h->fixed[cidx]) h->val[p] = (cidx == r && h->fixed[r]) ? 1.0 : 0.0;
        }
    return now() - t0;
}

static void spmv(const cref_t* h, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < h->ndof; r++) {
        double s = 0;
        for (long long p = h->rowptr[r]; p < h->rowptr[r + 1]; p++) s += h->val[p] * x[h->col[p]];
        y[r] = s;
    }
}

/* MatMult on the assembled matrix; returns seconds per application */
double cref_spmv(cref_t* h, const double* x, double* y, int reps) {
    spmv(h, x, y);
    const double t0 = now();
    for (int r = 0; r < reps; r++) spmv(h, x, y);
    return (now() - t0) / (reps > 0 ? reps : 1);
}

static double dot(long long n, const double* a, const double* b) {
    double s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (long long i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* Jacobi-PCG on K U = N.RHS, warm start from U.  Returns iterations; stops at ||r|| <= rtol ||b|| or maxit. */
int cref_pcg(cref_t* h, double* U, double rtol, int maxit, double* relres, double* seconds) {
    const long long n = h->ndof;
    double *r = malloc(n * 8), *z = malloc(n * 8), *p = malloc(n * 8), *q = malloc(n * 8), *b = malloc(n * 8), *dinv = malloc(n * 8);
    double* Uin = U;
    U = malloc(n * 8); /* working copy first-touched by the threads that use it (NUMA placement) */
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) {
        U[i] = Uin[i];
        b[i] = h->fixed[i] ? 0.0 : h->rhs[i]; /* RHS <- RHS.N, :542 */
        if (h->fixed[i]) U[i] = 0.0;
        double d = 1.0;
        for (long long pp = h->rowptr[i]; pp < h->rowptr[i + 1]; pp++)
            if (h->col[pp] == i) d = h->val[pp];
        dinv[i] = 1.0 / d;
    }
    const double bnorm = sqrt(dot(n, b, b));
    spmv(h, U, q);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) { r[i] = b[i] - q[i]; z[i] = r[i] * dinv[i]; p[i] = z[i]; }
    double rz = dot(n, r, z), rn = sqrt(dot(n, r, r));
    int it = 0;
    const double t0 = now(); /* the Krylov iterations only (set-up above is not part of an iteration) */
    while (rn > rtol * bnorm && it < maxit) {
        spmv(h, p, q);
        const double alpha = rz / dot(n, p, q);
        double rz2 = 0, rr = 0;
#pragma omp parallel for schedule(static) reduction(+ : rz2, rr)
        for (long long i = 0; i < n; i++) {
            U[i] += alpha * p[i];
            r[i] -= alpha * q[i];
            z[i] = r[i] * dinv[i];
            rz2 += r[i] * z[i];
            rr += r[i] * r[i];
        }
        const double beta = rz2 / rz;
        rz = rz2;
        rn = sqrt(rr);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        it++;
    }
    if (relres) *relres = bnorm > 0 ? rn / bnorm : 0;
    if (seconds) *seconds = now() - t0;
    memcpy(Uin, U, n * 8);
    free(U);
    free(r); free(z); free(p); free(q); free(b); free(dinv);
    return it;
}

/* objective and sensitivities, LinearElasticity.cc:403-437. Returns seconds. */
double cref_objective(cref_t* h, const double* U, const double* xphys, double Emin, double Emax, double penal,
                      double volfrac, double* fx, double* gx, double* dfdx) {
    const double t0 = now();
    double f = 0, vs = 0;
#pragma omp parallel for schedule(static) reduction(+ : f, vs)
    for (long long e = 0; e < h->nel; e++) {
        const int ei = e % h->nex, ej = (e / h->nex) % h->ney, ek = e / ((long long)h->nex * h->ney);
        double ue[24];
        for (int ln = 0; ln < 8; ln++) {
            const int a = (ln == 1 || ln == 2 || ln == 5 || ln == 6), b = (ln == 2 || ln == 3 || ln == 6 || ln == 7), g = ln >> 2;
            const long long m = nid(h, ei + a, ej + b, ek + g);
            for (int c = 0; c < 3; c++) ue[3 * ln + c] = U[3 * m + c];
        }
        double uKu = 0;
        for (int k = 0; k < 24; k++)
            for (int l = 0; l < 24; l++) uKu += ue[k] * h->KE[k * 24 + l] * ue[l];
        f += (Emin + pow(xphys[e], penal) * (Emax - Emin)) * uKu;
        if (dfdx) dfdx[e] = -1.0 * penal * pow(xphys[e], penal - 1) * (Emax - Emin) * uKu;
        vs += xphys[e];
    }
    *fx = f;
    *gx = vs / (double)h->nel - volfrac;
    return now() - t0;
}

/* =====================================================================================================
 * The reference's solver: FGMRES + PCMG (Galerkin) + GMRES/SOR smoothers.  LinearElasticity.cc:613-746.
 * ===================================================================================================== */

/* rows of one "process": thread t of nt owns whole nodes [n0,n1) in natural order (contiguous z-slabs, like the DMDA) */
static inline void my_rows(const lvl_t* L, int t, int nt, long long* r0, long long* r1) {
    *r0 = 3 * (L->nnode * t / nt);
    *r1 = 3 * (L->nnode * (t + 1) / nt);
}

static void lvl_spmv(const lvl_t* L, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->ndof; r++) {
        double s = 0;
        for (long long p = L->rowptr[r]; p < L->rowptr[r + 1]; p++) s += L->val[p] * x[L->col[p]];
        y[r] = s;
    }
}

static void lvl_diag(lvl_t* L) {
    if (!L->dinv) L->dinv = (double*)malloc(L->ndof * sizeof(double));
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->ndof; r++) {
        double d = 1.0;
        for (long long p = L->rowptr[r]; p < L->rowptr[r + 1]; p++)
            if (L->col[p] == r) d = L->val[p];
        L->dinv[r] = 1.0 / d;
    }
}

/* PCApply of PCSOR with PETSc's defaults (SOR_LOCAL_SYMMETRIC_SWEEP, omega 1, its 1, lits 1, zero initial guess):
 * z = (D+U)^-1 D (D+L)^-1 b on the diagonal block of every process -- off-process columns do not take part. */
static void lvl_sor(const lvl_t* L, const double* b, double* z) {
#pragma omp parallel
    {
        long long r0, r1;
        my_rows(L, omp_get_thread_num(), omp_get_num_threads(), &r0, &r1);
        for (long long r = r0; r < r1; r++) { /* forward: z = (D+L)^-1 b */
            double s = b[r];
            for (long long p = L->rowptr[r]; p < L->rowptr[r + 1]; p++) {
                const long long c = L->col[p];
                if (c >= r) break; /* columns are ascending inside a row */
                if (c >= r0) s -= L->val[p] * z[c];
            }
            z[r] = s * L->dinv[r];
        }
        for (long long r = r1 - 1; r >= r0; r--) { /* backward: z = (D+U)^-1 (b - L z) */
            double s = b[r];
            for (long long p = L->rowptr[r]; p < L->rowptr[r + 1]; p++) {
                const long long c = L->col[p];
                if (c != r && c >= r0 && c < r1) s -= L->val[p] * z[c];
            }
            z[r] = s * L->dinv[r];
        }
    }
}

static double vdot(long long n, const double* a, const double* b) {
    double s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (long long i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
static void vaxpy(long long n, double a, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] += a * x[i];
}
static void vscale_to(long long n, double a, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = a * x[i];
}

/* Left-preconditioned GMRES(m) with the SOR preconditioner on level L: A x = b.
 * maxit iterations at most; rtol < 0 means "no convergence test" (the PCMG smoothers, KSP_NORM_NONE); otherwise stop
 * when the preconditioned residual norm <= rtol * ||M^-1 b|| (KSP default for left preconditioning).
 * x_nonzero = 0: the initial guess is taken as zero (saves one MatMult, as KSP does).  Returns iterations. */
static int lvl_gmres_sor(lvl_t* L, const double* b, double* x, int m, int maxit, double rtol, int x_nonzero) {
    const long long n = L->ndof;
    double* V = L->V; /* (m+1) vectors */
    double* w = L->w;
    double* t = L->t;
    double H[32][31], cs[31], sn[31], g[32], y[31];
    int it = 0;
    double bnorm_pc = -1.0;
    while (it < maxit) {
        /* r = M^-1 (b - A x) */
        if (x_nonzero) {
            lvl_spmv(L, x, t);
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) t[i] = b[i] - t[i];
            lvl_sor(L, t, w);
        } else {
            lvl_sor(L, b, w);
        }
        double beta = sqrt(vdot(n, w, w));
        if (bnorm_pc < 0) {
            if (x_nonzero && rtol >= 0) { lvl_sor(L, b, t); bnorm_pc = sqrt(vdot(n, t, t)); }
            else bnorm_pc = beta;
        }
        if (beta == 0.0 || (rtol >= 0 && beta <= rtol * bnorm_pc)) break;
        vscale_to(n, 1.0 / beta, w, V);
        memset(g, 0, sizeof g);
        g[0] = beta;
        int k = 0;
        for (; k < m && it < maxit; k++, it++) {
            lvl_spmv(L, V + (long long)k * n, t);
            lvl_sor(L, t, w);
            /* classical Gram-Schmidt (KSPGMRESClassicalGramSchmidtOrthogonalization, PETSc's default) */
            for (int j = 0; j <= k; j++) H[j][k] = vdot(n, w, V + (long long)j * n);
            for (int j = 0; j <= k; j++) vaxpy(n, -H[j][k], V + (long long)j * n, w);
            const double hn = sqrt(vdot(n, w, w));
            H[k + 1][k] = hn;
            if (hn > 0) vscale_to(n, 1.0 / hn, w, V + (long long)(k + 1) * n);
            for (int j = 0; j < k; j++) { /* previous Givens rotations */
                const double a = H[j][k], bb = H[j + 1][k];
                H[j][k] = cs[j] * a + sn[j] * bb;
                H[j + 1][k] = -sn[j] * a + cs[j] * bb;
            }
            const double a = H[k][k], bb = H[k + 1][k], rr = hypot(a, bb);
            cs[k] = rr > 0 ? a / rr : 1.0; sn[k] = rr > 0 ? bb / rr : 0.0;
            H[k][k] = rr; H[k + 1][k] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            if (hn == 0.0 || (rtol >= 0 && fabs(g[k + 1]) <= rtol * bnorm_pc)) { k++; it++; break; }
        }
        for (int i = k - 1; i >= 0; i--) { /* back substitution, x += V y */
            double s2 = g[i];
            for (int j = i + 1; j < k; j++) s2 -= H[i][j] * y[j];
            y[i] = s2 / H[i][i];
        }
        if (!x_nonzero) { memset(x, 0, n * sizeof(double)); }
        for (int j = 0; j < k; j++) vaxpy(n, y[j], V + (long long)j * n, x);
        x_nonzero = 1;
        if (rtol >= 0 && fabs(g[k]) <= rtol * bnorm_pc) break;
    }
    return it;
}

/* DMDA Q1 interpolation (DMCreateInterpolation, LinearElasticity.cc:704), per dof: xf += P xc / bc = P^T rf */
static void mg_prolong_add(const lvl_t* F, const lvl_t* Cc, const double* xc, double* xf) {
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < F->nnode; n++) {
        const int i = n % F->nxn, j = (n / F->nxn) % F->nyn, k = n / ((long long)F->nxn * F->nyn);
        const int I0 = i >> 1, J0 = j >> 1, K0 = k >> 1, ni = 1 + (i & 1), nj = 1 + (j & 1), nk = 1 + (k & 1);
        const double w = 1.0 / (ni * nj * nk);
        double acc[3] = {0, 0, 0};
        for (int a = 0; a < nk; a++)
            for (int b = 0; b < nj; b++)
                for (int c = 0; c < ni; c++) {
                    const long long m = lnid(Cc, I0 + c, J0 + b, K0 + a);
                    acc[0] += xc[3 * m]; acc[1] += xc[3 * m + 1]; acc[2] += xc[3 * m + 2];
                }
        for (int c = 0; c < 3; c++) xf[3 * n + c] += w * acc[c];
    }
}
static void mg_restrict(const lvl_t* F, const lvl_t* Cc, const double* rf, double* bc) {
#pragma omp parallel for schedule(static)
    for (long long N = 0; N < Cc->nnode; N++) {
        const int I = N % Cc->nxn, J = (N / Cc->nxn) % Cc->nyn, K = N / ((long long)Cc->nxn * Cc->nyn);
        double acc[3] = {0, 0, 0};
        for (int dk = -1; dk <= 1; dk++) {
            const int z = 2 * K + dk;
            if (z < 0 || z >= F->nzn) continue;
            for (int dj = -1; dj <= 1; dj++) {
                const int y = 2 * J + dj;
                if (y < 0 || y >= F->nyn) continue;
                for (int di = -1; di <= 1; di++) {
                    const int x = 2 * I + di;
                    if (x < 0 || x >= F->nxn) continue;
                    const double w = (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0);
                    const long long m = lnid(F, x, y, z);
                    acc[0] += w * rf[3 * m]; acc[1] += w * rf[3 * m + 1]; acc[2] += w * rf[3 * m + 2];
                }
            }
        }
        for (int c = 0; c < 3; c++) bc[3 * N + c] = acc[c];
    }
}

/* Galerkin coarse operator A_c = P^T A_f P (PCMGSetGalerkin, LinearElasticity.cc:702; redone by KSPSetUp every design
 * iteration, :198-200).  Both matrices keep the box-stencil block structure, so the triple product is formed row by row:
 * for a coarse node I, every fine node n in the support of its hat function contributes w(n,I) * (A_f P)[n,:]. */
static inline int cbase(int i) { return ((i + 1) >> 1) - 1; } /* first coarse index reached from the stencil of fine index i */
static void mg_galerkin(const lvl_t* F, lvl_t* Cc) {
#pragma omp parallel
    {
        double* AP = (double*)malloc(27 * 9 * sizeof(double)); /* (A_f P)[n, 3x3x3 coarse slots][3x3] of one fine node */
#pragma omp for schedule(static)
        for (long long N = 0; N < Cc->nnode; N++) {
            const int I = N % Cc->nxn, J = (N / Cc->nxn) % Cc->nyn, K = N / ((long long)Cc->nxn * Cc->nyn);
            double acc[27][9];
            memset(acc, 0, sizeof acc);
            for (int dk = -1; dk <= 1; dk++) {
                const int fk = 2 * K + dk;
                if (fk < 0 || fk >= F->nzn) continue;
                for (int dj = -1; dj <= 1; dj++) {
                    const int fj = 2 * J + dj;
                    if (fj < 0 || fj >= F->nyn) continue;
                    for (int di = -1; di <= 1; di++) {
                        const int fi = 2 * I + di;
                        if (fi < 0 || fi >= F->nxn) continue;
                        const double wrow = (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0);
                        const long long n = lnid(F, fi, fj, fk);
                        const int bi = cbase(fi), bj = cbase(fj), bk = cbase(fk);
                        /* (A_f P)[n, :] */
                        memset(AP, 0, 27 * 9 * sizeof(double));
                        for (int ok = -1; ok <= 1; ok++)
                            for (int oj = -1; oj <= 1; oj++)
                                for (int oi = -1; oi <= 1; oi++) {
                                    const int blk = F->nbr_block[n * 27 + (ok + 1) * 9 + (oj + 1) * 3 + (oi + 1)];
                                    if (blk < 0) continue;
                                    const int mi = fi + oi, mj = fj + oj, mk = fk + ok;
                                    double a[9];
                                    for (int c = 0; c < 3; c++)
                                        for (int c2 = 0; c2 < 3; c2++) a[3 * c + c2] = F->val[F->rowptr[3 * n + c] + 3 * blk + c2];
                                    const int ci0 = mi >> 1, cj0 = mj >> 1, ck0 = mk >> 1, ni = 1 + (mi & 1), nj = 1 + (mj & 1), nk = 1 + (mk & 1);
                                    const double w = 1.0 / (ni * nj * nk);
                                    for (int a3 = 0; a3 < nk; a3++)
                                        for (int a2 = 0; a2 < nj; a2++)
                                            for (int a1 = 0; a1 < ni; a1++) {
                                                const int slot = ((ck0 + a3 - bk) * 3 + (cj0 + a2 - bj)) * 3 + (ci0 + a1 - bi);
                                                double* d = AP + 9 * slot;
                                                for (int q = 0; q < 9; q++) d[q] += w * a[q];
                                            }
                                }
                        /* acc[J - I] += wrow * AP[slot] */
                        for (int sk = 0; sk < 3; sk++)
                            for (int sj = 0; sj < 3; sj++)
                                for (int si = 0; si < 3; si++) {
                                    const int ck = bk + sk - K, cj = bj + sj - J, ci = bi + si - I; /* coarse offset of the column */
                                    if (ck < -1 || ck > 1 || cj < -1 || cj > 1 || ci < -1 || ci > 1) continue;
                                    const double* sAP = AP + 9 * ((sk * 3 + sj) * 3 + si);
                                    double* d = acc[(ck + 1) * 9 + (cj + 1) * 3 + (ci + 1)];
                                    for (int q = 0; q < 9; q++) d[q] += wrow * sAP[q];
                                }
                    }
                }
            }
            for (int o = 0; o < 27; o++) {
                const int blk = Cc->nbr_block[N * 27 + o];
                if (blk < 0) continue;
                for (int c = 0; c < 3; c++)
                    for (int c2 = 0; c2 < 3; c2++) Cc->val[Cc->rowptr[3 * N + c] + 3 * blk + c2] = acc[o][3 * c + c2];
            }
        }
        free(AP);
    }
}

static void lvl_alloc_work(lvl_t* L, int nbasis) {
    const long long n = L->ndof;
    if (!L->b) {
        L->b = (double*)calloc(n, 8); L->x = (double*)calloc(n, 8); L->r = (double*)calloc(n, 8);
        L->w = (double*)calloc(n, 8); L->t = (double*)calloc(n, 8);
        L->V = (double*)malloc((long long)nbasis * n * 8);
        /* first touch by the threads that will stream these vectors */
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++)
            for (int k = 0; k < nbasis; k++) L->V[(long long)k * n + i] = 0.0;
    }
}

/* KSPSetUp after an operator change: build the hierarchy on first use, redo the Galerkin products and SOR diagonals.
 * nlvls as in TopOpt.cc:116 (default 4).  Returns seconds. */
double cref_mg_setup(cref_t* h, int nlvls) {
    const double t0 = now();
    if (nlvls > CREF_MAXL) nlvls = CREF_MAXL;
    if (h->nl != nlvls) {
        int nx = h->nxn, ny = h->nyn, nz = h->nzn;
        for (int l = 1; l < nlvls; l++) {
            if (((nx - 1) | (ny - 1) | (nz - 1)) & 1) { fprintf(stderr, "cpu_ref: MESH DIMENSION NOT COMPATIBLE WITH NUMBER OF MULTIGRID LEVELS\n"); return -1.0; }
            nx = (nx - 1) / 2 + 1; ny = (ny - 1) / 2 + 1; nz = (nz - 1) / 2 + 1;
            if (l >= h->nl) lvl_pattern(&h->lv[l], nx, ny, nz);
        }
        h->nl = nlvls;
        for (int l = 0; l < nlvls; l++) lvl_alloc_work(&h->lv[l], l == nlvls - 1 ? 31 : 5);
    }
    for (int l = 1; l < h->nl; l++) mg_galerkin(&h->lv[l - 1], &h->lv[l]);
    for (int l = 0; l < h->nl; l++) lvl_diag(&h->lv[l]);
    return now() - t0;
}

/* PCApply of PCMG, multiplicative V-cycle: x_l = MG(b_l).  Levels are numbered fine = 0 here (PETSc numbers them the
 * other way round, LinearElasticity.cc:684-689). */
static void mg_vcycle(cref_t* h, int l) {
    lvl_t* L = &h->lv[l];
    if (l == h->nl - 1) { /* coarse solve: GMRES(30) + SOR, rtol 1e-8, <= 30 its (:720-731) */
        h->mg_coarse_its += lvl_gmres_sor(L, L->b, L->x, 30, 30, 1.0e-8, 0);
        return;
    }
    lvl_gmres_sor(L, L->b, L->x, 4, 4, -1.0, 0); /* pre-smoothing: GMRES(4) x 4 its + SOR (:734-746) */
    h->mg_smooth_calls++;
    lvl_spmv(L, L->x, L->r);
    const long long n = L->ndof;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) L->r[i] = L->b[i] - L->r[i];
    lvl_t* Cc = &h->lv[l + 1];
    mg_restrict(L, Cc, L->r, Cc->b);
    mg_vcycle(h, l + 1);
    mg_prolong_add(L, Cc, Cc->x, L->x);
    lvl_gmres_sor(L, L->b, L->x, 4, 4, -1.0, 1); /* post-smoothing, nonzero initial guess */
    h->mg_smooth_calls++;
}

static void cref_free_krylov(cref_t* h) {
    for (int i = 0; i < h->fg_nv; i++) free(h->fgV[i]);
    for (int i = 0; i < h->fg_nz; i++) free(h->fgZ[i]);
    free(h->fgV); free(h->fgZ);
    h->fgV = h->fgZ = NULL; h->fg_nv = h->fg_nz = 0;
}
static double* fg_vec(double*** arr, int* n, int idx, long long len) {
    if (!*arr) *arr = (double**)calloc(256, sizeof(double*));
    while (*n <= idx) {
        double* v = (double*)malloc(len * 8);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < len; i++) v[i] = 0.0;
        (*arr)[(*n)++] = v;
    }
    return (*arr)[idx];
}

/* KSPSolve(ksp, RHS, U) of SolveState (LinearElasticity.cc:204): FGMRES(restart 100), right preconditioning with the
 * V-cycle above, nonzero initial guess, ||r|| <= max(rtol ||b||, atol), divergence at dtol ||b||, <= maxits.
 * cref_mg_setup must have been called for the current matrix.  Returns iterations. */
int cref_fgmres_mg(cref_t* h, double* U, double rtol, int maxits, int restart, double* relres, double* seconds) {
    const double t0 = now();
    const long long n = h->ndof;
    lvl_t* L0 = &h->lv[0];
    if (restart > 100) restart = 100;
    const double atol = 1.0e-50, dtol = 1.0e5;
    static double H[101][100];
    double cs[100], sn[100], g[101], y[100];
    double* bb = (double*)malloc(n * 8);
    double* w = (double*)malloc(n * 8);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) {
        bb[i] = h->fixed[i] ? 0.0 : h->rhs[i]; /* RHS <- RHS.N, :542 */
        w[i] = 0.0;
    }
    const double* b = bb;
    const double bnorm = sqrt(vdot(n, b, b));
    const double tol = fmax(rtol * bnorm, atol);
    int it = 0;
    double rn = 0;
    int done = 0;
    while (!done) {
        lvl_spmv(L0, U, w);
        double* v0 = fg_vec(&h->fgV, &h->fg_nv, 0, n);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++) v0[i] = b[i] - w[i];
        rn = sqrt(vdot(n, v0, v0));
        if (rn <= tol || rn > dtol * bnorm || it >= maxits) break;
        vscale_to(n, 1.0 / rn, v0, v0);
        memset(g, 0, sizeof g);
        g[0] = rn;
        int k = 0;
        for (; k < restart && it < maxits; ) {
            double* vk = h->fgV[k];
            double* zk = fg_vec(&h->fgZ, &h->fg_nz, k, n);
            /* z_k = MG(v_k) */
            memcpy(L0->b, vk, n * 8);
            mg_vcycle(h, 0);
            memcpy(zk, L0->x, n * 8);
            lvl_spmv(L0, zk, w);
            for (int j = 0; j <= k; j++) H[j][k] = vdot(n, w, h->fgV[j]);
            for (int j = 0; j <= k; j++) vaxpy(n, -H[j][k], h->fgV[j], w);
            const double hn = sqrt(vdot(n, w, w));
            double* vn = fg_vec(&h->fgV, &h->fg_nv, k + 1, n);
            if (hn > 0) vscale_to(n, 1.0 / hn, w, vn);
            H[k + 1][k] = hn;
            for (int j = 0; j < k; j++) {
                const double a = H[j][k], c2 = H[j + 1][k];
                H[j][k] = cs[j] * a + sn[j] * c2;
                H[j + 1][k] = -sn[j] * a + cs[j] * c2;
            }
            const double a = H[k][k], c2 = H[k + 1][k], rr = hypot(a, c2);
            cs[k] = rr > 0 ? a / rr : 1.0; sn[k] = rr > 0 ? c2 / rr : 0.0;
            H[k][k] = rr; H[k + 1][k] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            k++; it++;
            rn = fabs(g[k]);
            if (rn <= tol || rn > dtol * bnorm || hn == 0.0) { done = 1; break; }
        }
        if (it >= maxits) done = 1;
        for (int i = k - 1; i >= 0; i--) {
            double s2 = g[i];
            for (int j = i + 1; j < k; j++) s2 -= H[i][j] * y[j];
            y[i] = s2 / H[i][i];
        }
        for (int j = 0; j < k; j++) vaxpy(n, y[j], h->fgZ[j], U);
    }
    if (relres) *relres = bnorm > 0 ? rn / bnorm : 0.0;
    free(bb); free(w);
    if (seconds) *seconds = now() - t0;
    return it;
}

void cref_mg_stats(cref_t* h, long long* smooth_calls, long long* coarse_its) {
    *smooth_calls = h->mg_smooth_calls; *coarse_its = h->mg_coarse_its;
}
/* a level's matrix as CSR (tests: Galerkin product vs scipy P^T A P) */
long long cref_level_nnz(cref_t* h, int l) { return l < h->nl ? h->lv[l].nnz : -1; }
long long cref_level_ndof(cref_t* h, int l) { return l < h->nl ? h->lv[l].ndof : -1; }
void cref_level_csr(cref_t* h, int l, long long* rowptr, int* col, double* val) {
    const lvl_t* L = &h->lv[l];
    memcpy(rowptr, L->rowptr, (L->ndof + 1) * sizeof(long long));
    memcpy(col, L->col, L->nnz * sizeof(int));
    memcpy(val, L->val, L->nnz * sizeof(double));
}
/* one V-cycle on a given right-hand side (tests) */
void cref_mg_apply(cref_t* h, const double* b, double* x) {
    memcpy(h->lv[0].b, b, h->ndof * 8);
    mg_vcycle(h, 0);
    memcpy(x, h->lv[0].x, h->ndof * 8);
}

/* =====================================================================================================
 * Cone filter as the reference applies it (Filter.cc): out = (H in) with H_ij = max(0, R - dist(c_i, c_j)) over the
 * clipped box +-s (Filter.cc:404-440).  mode 0: out = H in / Hs (FilterProject type 1, :66-71);
 * mode 1: out = H (in / Hs) (Gradients type 1, :178-192); Hs = H 1 (:445-448).  Returns seconds.
 * ===================================================================================================== */
double cref_cone_filter(cref_t* h, double rmin, int mode, const double* in, double* out) {
    const double t0 = now();
    const int nex = h->nex, ney = h->ney, nez = h->nez;
    int s = 0;
    { /* ElemConn, Filter.cc:326-332 */
        int sx = (int)ceil(rmin / h->dx) - 1, sy = (int)ceil(rmin / h->dy) - 1, sz = (int)ceil(rmin / h->dz) - 1;
        s = sx > sy ? sx : sy; s = s > sz ? s : sz;
        int cap = (nex - 1) / 2; if ((ney - 1) / 2 < cap) cap = (ney - 1) / 2; if ((nez - 1) / 2 < cap) cap = (nez - 1) / 2;
        if (s > cap) s = cap;
        if (s < 0) s = 0;
    }
    const int W = 2 * s + 1;
    double* wt = (double*)malloc((size_t)W * W * W * sizeof(double));
    for (int dk = -s; dk <= s; dk++)
        for (int dj = -s; dj <= s; dj++)
            for (int di = -s; di <= s; di++) {
                const double d = sqrt((di * h->dx) * (di * h->dx) + (dj * h->dy) * (dj * h->dy) + (dk * h->dz) * (dk * h->dz));
                wt[((dk + s) * W + (dj + s)) * W + (di + s)] = d < rmin ? rmin - d : 0.0;
            }
    const long long nel = h->nel;
    double* tmp = (double*)malloc(nel * 8);
    double* Hs = (double*)malloc(nel * 8);
    for (int pass = 0; pass < 2; pass++) { /* pass 0: Hs = H 1 ; pass 1: the product */
#pragma omp parallel for schedule(static)
        for (long long e = 0; e < nel; e++) {
            const int i = e % nex, j = (e / nex) % ney, k = e / ((long long)nex * ney);
            double acc = 0;
            for (int dk = -s; dk <= s; dk++) {
                const int kk = k + dk;
                if (kk < 0 || kk >= nez) continue;
                for (int dj = -s; dj <= s; dj++) {
                    const int jj = j + dj;
                    if (jj < 0 || jj >= ney) continue;
                    const double* wrow = wt + ((dk + s) * W + (dj + s)) * W + s;
                    const long long base = (long long)nex * (jj + (long long)ney * kk);
                    const int lo = i - s < 0 ? -i : -s, hi = i + s >= nex ? nex - 1 - i : s;
                    if (pass == 0) for (int di = lo; di <= hi; di++) acc += wrow[di];
                    else for (int di = lo; di <= hi; di++) acc += wrow[di] * tmp[base + i + di];
                }
            }
            if (pass == 0) Hs[e] = acc; else out[e] = mode == 0 ? acc / Hs[e] : acc;
        }
        if (pass == 0) {
#pragma omp parallel for schedule(static)
            for (long long e = 0; e < nel; e++) tmp[e] = mode == 1 ? in[e] / Hs[e] : in[e];
        }
    }
    free(wt); free(tmp); free(Hs);
    return now() - t0;
}
