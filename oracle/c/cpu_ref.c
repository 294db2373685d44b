/* cpu_ref.c -- C/OpenMP restatement of the reference's CPU state-solve path (TEST/BASELINE
 * INFRASTRUCTURE ONLY: used by tests/, bench.py's cpu_baseline / --impl reference legs, never by
 * the product path).
 *
 * What it restates (paths relative to the reference checkout):
 *   - Hex8Isoparametric, LinearElasticity.cc:841-998 (2x2x2 Gauss, B^T C B)
 *   - AssembleStiffnessMatrix, LinearElasticity.cc:487-549: K = sum_e E_e KE into an AIJ/CSR
 *     matrix with the DMDA box-stencil pattern (81 nnz per interior row), then
 *     K <- N K N + (I-N)
 *   - the Krylov solve of SolveState (LinearElasticity.cc:182-223) on the assembled matrix.
 *     PETSc (FGMRES + PCMG) is not available in this environment; the Krylov method restated here
 *     is Jacobi-preconditioned CG on the same assembled matrix, i.e. the same algorithm as the GPU
 *     arm's cfg2 solver, so both arms do identical numerical work per iteration.
 *   - objective / sensitivities, LinearElasticity.cc:403-437.
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

typedef struct {
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
} cref_t;

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

cref_t* cref_create(int nxn, int nyn, int nzn, double dx, double dy, double dz, double nu, int nthreads) {
    cref_t* h = (cref_t*)calloc(1, sizeof(cref_t));
    if (nthreads > 0) omp_set_num_threads(nthreads);
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
    h->rowptr = (long long*)malloc((h->ndof + 1) * sizeof(long long));
    h->nbr_block = (int*)malloc(h->nnode * 27 * sizeof(int));
    long long* cnt = (long long*)malloc((h->nnode + 1) * sizeof(long long));
    cnt[0] = 0;
    for (int k = 0; k < nzn; k++)
        for (int j = 0; j < nyn; j++)
            for (int i = 0; i < nxn; i++) {
                long long n = nid(h, i, j, k);
                int c = 0;
                for (int dk = -1; dk <= 1; dk++)
                    for (int dj = -1; dj <= 1; dj++)
                        for (int di = -1; di <= 1; di++) {
                            int ok = i + di >= 0 && i + di < nxn && j + dj >= 0 && j + dj < nyn && k + dk >= 0 && k + dk < nzn;
                            h->nbr_block[n * 27 + (dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] = ok ? c : -1;
                            c += ok;
                        }
                cnt[n + 1] = cnt[n] + c;
            }
    h->nnz = 9 * cnt[h->nnode];
    h->col = (int*)malloc(h->nnz * sizeof(int));
    h->val = (double*)malloc(h->nnz * sizeof(double));
    h->E = (double*)malloc(h->nel * sizeof(double));
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < h->nnode; n++) {
        const int i = n % nxn, j = (n / nxn) % nyn, k = n / ((long long)nxn * nyn);
        const long long nb = cnt[n + 1] - cnt[n];
        for (int c = 0; c < 3; c++) h->rowptr[3 * n + c] = 9 * cnt[n] + c * 3 * nb;
        for (int dk = -1; dk <= 1; dk++)
            for (int dj = -1; dj <= 1; dj++)
                for (int di = -1; di <= 1; di++) {
                    int b = h->nbr_block[n * 27 + (dk + 1) * 9 + (dj + 1) * 3 + (di + 1)];
                    if (b < 0) continue;
                    long long m = nid(h, i + di, j + dj, k + dk);
                    for (int c = 0; c < 3; c++)
                        for (int c2 = 0; c2 < 3; c2++) h->col[h->rowptr[3 * n + c] + 3 * b + c2] = (int)(3 * m + c2);
                }
    }
    h->rowptr[h->ndof] = h->nnz;
    free(cnt);
    return h;
}

void cref_destroy(cref_t* h) {
    if (!h) return;
    free(h->fixed); free(h->rhs); free(h->rowptr); free(h->col); free(h->val); free(h->nbr_block); free(h->E);
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
