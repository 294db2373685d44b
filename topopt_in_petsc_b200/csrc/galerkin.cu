// galerkin.cu -- Galerkin coarse operators for the multigrid preconditioner, as assembled 27-point block stencils.
//
// The reference builds every coarse operator as R A P (PCMGSetGalerkin, LinearElasticity.cc:702) and redoes the products
// every design iteration (:198-200).  Round 1 re-discretised the coarse levels with the 8-child mean of the SIMP modulus --
// cheap, but at 0/1 designs (contrast 1e9) thin members are smeared on the coarse grids and the PCG iteration count grew
// from 19 to 83 over 20 design iterations.  Here the levels from `gf` down ARE Galerkin products:
//   * level gf from its matrix-free parent: K_C = sum over the 8 children of E_child * (Q_child^T KE Q_child); the eight
//     24x24 matrices Q^T KE Q depend on the geometry only (host, once), so the product costs 64 FMA per stencil entry;
//   * deeper levels from the assembled parent stencil (P^T A P row by row, one CTA per coarse node).
// Storage: 27 neighbour blocks of 3x3 per node, struct-of-arrays A[k][node] (k = 9*offset + 3*c + c2) so that the
// one-thread-per-node SpMV reads every coefficient stream coalesced; fixed dofs carry identity rows/columns.
// The levels handled here are small (replicated on every rank, gmg.cu): the finest one is 1/64 of the design grid or less.
#include "elast.cuh"
#include "galerkin.cuh"
#include "vecops.cuh"

namespace tob {

// ---- host: Q_child^T KE Q_child for the 8 children (binary local node index a + 2b + 4g on both levels) ------------------
static const int kBin2RefG[8] = {0, 1, 3, 2, 4, 5, 7, 6};
void galerkin_child_matrices(const double* KE_fine, double* M /* [8][24][24] */) {
    auto w1 = [](int pos, int l) { return pos == 1 ? 0.5 : ((pos == 0) == (l == 0) ? 1.0 : 0.0); };
    for (int ch = 0; ch < 8; ch++) {
        const int ca = ch & 1, cb = (ch >> 1) & 1, cg = ch >> 2;
        double Q[8][8];  // fine local node f <- coarse local node l (scalar; the same for the three components)
        for (int f = 0; f < 8; f++)
            for (int l = 0; l < 8; l++)
                Q[f][l] = w1(ca + (f & 1), l & 1) * w1(cb + ((f >> 1) & 1), (l >> 1) & 1) * w1(cg + (f >> 2), l >> 2);
        for (int l = 0; l < 8; l++)
            for (int c = 0; c < 3; c++)
                for (int l2 = 0; l2 < 8; l2++)
                    for (int c2 = 0; c2 < 3; c2++) {
                        double s = 0.0;
                        for (int f = 0; f < 8; f++) {
                            if (Q[f][l] == 0.0) continue;
                            for (int f2 = 0; f2 < 8; f2++)
                                s += Q[f][l] * KE_fine[(3 * kBin2RefG[f] + c) * 24 + 3 * kBin2RefG[f2] + c2] * Q[f2][l2];
                        }
                        M[(ch * 24 + 3 * l + c) * 24 + 3 * l2 + c2] = s;
                    }
    }
}

// ---- assembly from a matrix-free parent ----------------------------------------------------------------------------------
// one CTA per coarse node, thread t < 243 owns entry (o, c, c2) of its stencil row block
template <typename T>
__global__ void __launch_bounds__(256)
k_galerkin_mf(const double* __restrict__ Ef, const double* __restrict__ M, const unsigned char* __restrict__ maskc, T* __restrict__ A,
              int ncx, int ncy, int ncz, int pxc, int pmc, long long msc, int nfex, int nfey, int nfez, int pef, long long esf,
              long long nnp) {
    const int I = blockIdx.x, J = blockIdx.y, K = blockIdx.z;
    const int t = threadIdx.x;
    if (t >= 243) return;
    const int o = t / 9, q = t % 9, c = q / 3, c2 = q % 3;
    const int di = o % 3 - 1, dj = (o / 3) % 3 - 1, dk = o / 9 - 1;
    const int I2 = I + di, J2 = J + dj, K2 = K + dk;
    const long long n = ((long long)K * ncy + J) * pxc + I;
    double acc = 0.0;
    const bool inb = I2 >= 0 && I2 < ncx && J2 >= 0 && J2 < ncy && K2 >= 0 && K2 < ncz;
    if (inb) {
        for (int eg = -1; eg <= 0; eg++)
            for (int eb = -1; eb <= 0; eb++)
                for (int ea = -1; ea <= 0; ea++) {
                    // coarse element (I+ea, J+eb, K+eg); both nodes must belong to it
                    const int la = -ea, lb = -eb, lg = -eg;              // local position of the row node
                    const int ma = di - ea, mb = dj - eb, mg = dk - eg;  // local position of the column node
                    if (ma < 0 || ma > 1 || mb < 0 || mb > 1 || mg < 0 || mg > 1) continue;
                    const int ei = I + ea, ej = J + eb, ek = K + eg;
                    if (ei < 0 || ej < 0 || ek < 0 || 2 * ei >= nfex || 2 * ej >= nfey || 2 * ek >= nfez) continue;
                    const int row = 3 * (la + 2 * lb + 4 * lg) + c, col = 3 * (ma + 2 * mb + 4 * mg) + c2;
#pragma unroll
                    for (int ch = 0; ch < 8; ch++) {
                        const double E = Ef[(long long)(2 * ek + (ch >> 2)) * esf + (long long)(2 * ej + ((ch >> 1) & 1)) * pef + 2 * ei + (ch & 1)];
                        acc = fma(E, __ldg(&M[(ch * 24 + row) * 24 + col]), acc);
                    }
                }
        // Dirichlet: N A N + (I - N) with the injected coarse mask
        const unsigned char mr = maskc[(long long)K * msc + (long long)J * pmc + I];
        const unsigned char mc = maskc[(long long)K2 * msc + (long long)J2 * pmc + I2];
        const bool fr = (mr >> c) & 1, fc = (mc >> c2) & 1;
        if (fr || fc) acc = (o == 13 && c == c2 && fr) ? 1.0 : 0.0;
    }
    A[(long long)t * nnp + n] = (T)acc;
}

// ---- assembly from an assembled parent stencil: A_c = P^T A_f P ------------------------------------------------------------
__device__ __forceinline__ double q1w(int fine, int coarse) {  // 1-D interpolation weight of coarse node `coarse` at fine node `fine`
    const int d = fine - 2 * coarse;
    return d == 0 ? 1.0 : ((d == 1 || d == -1) ? 0.5 : 0.0);
}
template <typename T>
__global__ void __launch_bounds__(256)
k_galerkin_st(const T* __restrict__ Af, const unsigned char* __restrict__ maskc, T* __restrict__ Ac, int nfx, int nfy, int nfz,
              int pxf, long long nnpf, int ncx, int ncy, int ncz, int pxc, int pmc, long long msc, long long nnpc) {
    const int I = blockIdx.x, J = blockIdx.y, K = blockIdx.z;
    const int t = threadIdx.x;
    if (t >= 243) return;
    const int o = t / 9, q = t % 9, c = q / 3, c2 = q % 3;
    const int di = o % 3 - 1, dj = (o / 3) % 3 - 1, dk = o / 9 - 1;
    const int I2 = I + di, J2 = J + dj, K2 = K + dk;
    const long long n = ((long long)K * ncy + J) * pxc + I;
    double acc = 0.0;
    const bool inb = I2 >= 0 && I2 < ncx && J2 >= 0 && J2 < ncy && K2 >= 0 && K2 < ncz;
    if (inb) {
        for (int rk = -1; rk <= 1; rk++) {
            const int fk = 2 * K + rk;
            if (fk < 0 || fk >= nfz) continue;
            for (int rj = -1; rj <= 1; rj++) {
                const int fj = 2 * J + rj;
                if (fj < 0 || fj >= nfy) continue;
                for (int ri = -1; ri <= 1; ri++) {
                    const int fi = 2 * I + ri;
                    if (fi < 0 || fi >= nfx) continue;
                    const double wrow = (ri ? 0.5 : 1.0) * (rj ? 0.5 : 1.0) * (rk ? 0.5 : 1.0);
                    const long long nf = ((long long)fk * nfy + fj) * pxf + fi;
                    double s = 0.0;
                    for (int ok = -1; ok <= 1; ok++) {
                        const double wz = q1w(fk + ok, K2);
                        if (wz == 0.0 || fk + ok < 0 || fk + ok >= nfz) continue;
                        for (int oj = -1; oj <= 1; oj++) {
                            const double wy = q1w(fj + oj, J2);
                            if (wy == 0.0 || fj + oj < 0 || fj + oj >= nfy) continue;
                            for (int oi = -1; oi <= 1; oi++) {
                                const double wx = q1w(fi + oi, I2);
                                if (wx == 0.0 || fi + oi < 0 || fi + oi >= nfx) continue;
                                const int of = ((ok + 1) * 3 + (oj + 1)) * 3 + (oi + 1);
                                s = fma(wx * wy * wz, (double)Af[(long long)(of * 9 + q) * nnpf + nf], s);
                            }
                        }
                    }
                    acc = fma(wrow, s, acc);
                }
            }
        }
        const unsigned char mr = maskc[(long long)K * msc + (long long)J * pmc + I];
        const unsigned char mc = maskc[(long long)K2 * msc + (long long)J2 * pmc + I2];
        const bool fr = (mr >> c) & 1, fc = (mc >> c2) & 1;
        if (fr || fc) acc = (o == 13 && c == c2 && fr) ? 1.0 : 0.0;
    }
    Ac[(long long)t * nnpc + n] = (T)acc;
}

// ---- y = A x with the epilogues of the matrix-free kernel (one thread per node) -------------------------------------------
// coef: device array {s0, s1} pairs of the Chebyshev recurrence (EPI_CHEB), indexed by `step`
template <int EPI, typename T>
__global__ void __launch_bounds__(128)
k_stencil_apply(const T* __restrict__ A, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ a0,
                const double* __restrict__ a1, double* __restrict__ o1, const double* __restrict__ coef, int step, int nx, int ny, int nz,
                int px, long long nstride, long long pstride, long long nnp, double* __restrict__ partials) {
    __shared__ double red[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    double part = 0.0;
    if (i < nx) {
        const long long n = ((long long)k * ny + j) * px + i;
        double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
        for (int dk = -1; dk <= 1; dk++) {
            const int kk = k + dk;
            if (kk < 0 || kk >= nz) continue;
#pragma unroll 1
            for (int dj = -1; dj <= 1; dj++) {
                const int jj = j + dj;
                if (jj < 0 || jj >= ny) continue;
#pragma unroll
                for (int di = -1; di <= 1; di++) {
                    const int ii = i + di;
                    if (ii < 0 || ii >= nx) continue;
                    const int o = ((dk + 1) * 3 + (dj + 1)) * 3 + (di + 1);
                    const long long xo = (long long)kk * pstride + (long long)jj * px + ii;
                    const double x0 = x[xo], x1 = x[xo + nstride], x2 = x[xo + 2 * nstride];
                    const T* Ao = A + (long long)(o * 9) * nnp + n;
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        acc[c] = fma((double)Ao[(long long)(3 * c) * nnp], x0, fma((double)Ao[(long long)(3 * c + 1) * nnp], x1, fma((double)Ao[(long long)(3 * c + 2) * nnp], x2, acc[c])));
                }
            }
        }
        const long long v0 = (long long)k * pstride + (long long)j * px + i;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const long long idx = v0 + c * nstride;
            const double au = acc[c];
            if (EPI == EPI_Y) {
                y[idx] = au;
            } else if (EPI == EPI_DOT) {
                y[idx] = au;
                part = fma(x[idx], au, part);
            } else if (EPI == EPI_RESID) {
                const double r = a0[idx] - au;
                y[idx] = r;
                part = fma(r, r, part);
            } else {  // EPI_CHEB: d = s0 d + s1 dinv (b - A x) ; y = x + d
                const double r = a0[idx] - au;
                const double dn = coef[2 * step] * o1[idx] + coef[2 * step + 1] * a1[idx] * r;
                o1[idx] = dn;
                y[idx] = x[idx] + dn;
            }
        }
    }
    if (EPI == EPI_DOT || EPI == EPI_RESID) {
        part = block_sum(part, red);
        if (threadIdx.x == 0) partials[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = part;
    }
}

int stencil_apply(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, int epi, const double* x, double* y, const double* a0,
                  const double* a1, double* o1, int step, int* nparts) {
    dim3 grid(div_up(g.nxn, 128), g.nyn, g.nzn);
    const long long nb = (long long)grid.x * grid.y * grid.z;
    if (nb > kMaxPartials) TOB_FAIL("stencil level too large for the partial-sum buffer");
#define TOB_ST_LAUNCH(E)                                                                                                              \
    do {                                                                                                                              \
        if (S.fp32) k_stencil_apply<E, float><<<grid, 128, 0, c->stream>>>((const float*)S.A, x, y, a0, a1, o1, S.coef, step, g.nxn, g.nyn, g.nzn, g.px, g.nstride, g.pstride, S.nnp, c->d_partials); \
        else k_stencil_apply<E, double><<<grid, 128, 0, c->stream>>>((const double*)S.A, x, y, a0, a1, o1, S.coef, step, g.nxn, g.nyn, g.nzn, g.px, g.nstride, g.pstride, S.nnp, c->d_partials); \
    } while (0)
    switch (epi) {
        case EPI_Y: TOB_ST_LAUNCH(EPI_Y); break;
        case EPI_DOT: TOB_ST_LAUNCH(EPI_DOT); break;
        case EPI_RESID: TOB_ST_LAUNCH(EPI_RESID); break;
        default: TOB_ST_LAUNCH(EPI_CHEB); break;
    }
#undef TOB_ST_LAUNCH
    c->launches++;
    TOB_KERNEL_CHECK();
    if (nparts) *nparts = (epi == EPI_DOT || epi == EPI_RESID) ? (int)nb : 0;
    return 0;
}

// ---- diagonal, eigenvalue estimate, Chebyshev coefficients -----------------------------------------------------------------
template <typename T>
__global__ void k_stencil_diag(const T* __restrict__ A, double* __restrict__ dinv, int nx, int ny, int nz, int px, long long nstride,
                               long long pstride, long long nnp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= nx) return;
    const long long n = ((long long)k * ny + j) * px + i;
    for (int c = 0; c < 3; c++) {
        const double d = (double)A[(long long)(13 * 9 + 4 * c) * nnp + n];
        dinv[(long long)k * pstride + c * nstride + (long long)j * px + i] = d != 0.0 ? 1.0 / d : 1.0;
    }
}
// v <- D^-1 w / ||D^-1 w|| needs two passes; fused as: t = dinv .* w, partial |t|^2 (then scaled by the next kernel)
__global__ void __launch_bounds__(kFlatThreads) k_pw_scale_norm(double* __restrict__ t, const double* __restrict__ w, const double* __restrict__ dinv,
                                                                long long n, double* __restrict__ partials) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = dinv[i] * w[i];
        t[i] = v;
        s = fma(v, v, s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// one CTA: nrm = sqrt(sum partials); lam = nrm (since ||v|| = 1 before the step); v = t / nrm
__global__ void __launch_bounds__(256) k_power_finish(double* __restrict__ v, const double* __restrict__ t, long long n, const double* __restrict__ partials,
                                                      int nb, double* __restrict__ lam_out) {
    __shared__ double red[32];
    __shared__ double bc;
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partials[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) { bc = sqrt(s); *lam_out = bc; }
    __syncthreads();
    const double inv = bc > 0.0 ? 1.0 / bc : 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) v[i] = t[i] * inv;
}
// Chebyshev coefficients for `deg` steps targeting [lam/ratio, 1.1 lam] (same recurrence as gmg.cu's host version)
__global__ void k_cheb_coef(const double* __restrict__ lam_est, double ratio, int deg, double* __restrict__ coef) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double lam = 1.15 * *lam_est;  // the power iteration approaches lambda_max from below
    const double lmax = 1.02 * lam, lmin = lam / ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    coef[0] = 0.0; coef[1] = 1.0 / theta;
    for (int s = 1; s < deg; s++) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        coef[2 * s] = rho_new * rho;
        coef[2 * s + 1] = 2.0 * rho_new / delta;
        rho = rho_new;
    }
}
// first Chebyshev step from a zero iterate with the coefficient read from the device: d = coef[1] dinv b ; x = d
__global__ void __launch_bounds__(kFlatThreads) k_cheb_first_dev(double* __restrict__ x, double* __restrict__ d, const double* __restrict__ b,
                                                                 const double* __restrict__ dinv, const double* __restrict__ coef, long long n) {
    const double s1 = coef[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = s1 * dinv[i] * b[i];
        d[i] = v;
        x[i] = v;
    }
}
int stencil_cheb_first(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, double* x, double* d, const double* b, const double* dinv) {
    k_cheb_first_dev<<<flat_grid(c, g.nvec), kFlatThreads, 0, c->stream>>>(x, d, b, dinv, S.coef, g.nvec);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// ---- coarsest level: Jacobi-PCG to a relative residual inside ONE CTA (no launches, no host) ----------------------------
// x = A^-1 b approximately: stops at ||r|| <= rtol ||b|| or maxit.  Work vectors r, p, q in global memory (L1-resident).
template <typename T>
__global__ void __launch_bounds__(1024)
k_coarse_cg(const T* __restrict__ A, const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ x,
            double* __restrict__ r, double* __restrict__ p, double* __restrict__ q, int nx, int ny, int nz, int px, long long nstride,
            long long pstride, long long nnp, double rtol, int maxit, int* __restrict__ its_out) {
    __shared__ double red[32];
    __shared__ double bc[2];
    const int nn = nx * ny * nz;
    auto vidx = [&](int node, int c) {
        const int i = node % nx, j = (node / nx) % ny, k = node / (nx * ny);
        return (long long)k * pstride + c * nstride + (long long)j * px + i;
    };
    auto bsum = [&](double v) {
        v = block_sum(v, red);
        if (threadIdx.x == 0) bc[0] = v;
        __syncthreads();
        v = bc[0];
        __syncthreads();
        return v;
    };
    // x = 0, r = b, p = z = dinv r
    double rz = 0.0, rr = 0.0;
    for (int d = threadIdx.x; d < 3 * nn; d += blockDim.x) {
        const long long id = vidx(d / 3, d % 3);
        const double rv = b[id], zv = dinv[id] * rv;
        x[id] = 0.0; r[id] = rv; p[id] = zv;
        rz = fma(rv, zv, rz); rr = fma(rv, rv, rr);
    }
    rz = bsum(rz);
    rr = bsum(rr);
    const double stop = rtol * rtol * rr;
    int it = 0;
    while (it < maxit && rr > stop && rr > 0.0) {
        // q = A p, pq
        double pq = 0.0;
        for (int d = threadIdx.x; d < 3 * nn; d += blockDim.x) {
            const int node = d / 3, c = d % 3;
            const int i = node % nx, j = (node / nx) % ny, k = node / (nx * ny);
            const long long n = ((long long)k * ny + j) * px + i;
            double acc = 0.0;
            for (int dk = -1; dk <= 1; dk++) {
                const int kk = k + dk;
                if (kk < 0 || kk >= nz) continue;
                for (int dj = -1; dj <= 1; dj++) {
                    const int jj = j + dj;
                    if (jj < 0 || jj >= ny) continue;
                    for (int di = -1; di <= 1; di++) {
                        const int ii = i + di;
                        if (ii < 0 || ii >= nx) continue;
                        const int o = ((dk + 1) * 3 + (dj + 1)) * 3 + (di + 1);
                        const long long xo = (long long)kk * pstride + (long long)jj * px + ii;
                        const T* Ao = A + (long long)(o * 9 + 3 * c) * nnp + n;
                        acc = fma((double)Ao[0], p[xo], fma((double)Ao[nnp], p[xo + nstride], fma((double)Ao[2 * nnp], p[xo + 2 * nstride], acc)));
                    }
                }
            }
            const long long id = (long long)k * pstride + c * nstride + (long long)j * px + i;
            q[id] = acc;
            pq = fma(p[id], acc, pq);
        }
        pq = bsum(pq);
        const double alpha = rz / pq;
        double rz2 = 0.0, rr2 = 0.0;
        for (int d = threadIdx.x; d < 3 * nn; d += blockDim.x) {
            const long long id = vidx(d / 3, d % 3);
            x[id] = fma(alpha, p[id], x[id]);
            const double rv = fma(-alpha, q[id], r[id]);
            r[id] = rv;
            const double zv = dinv[id] * rv;
            rz2 = fma(rv, zv, rz2); rr2 = fma(rv, rv, rr2);
        }
        rz2 = bsum(rz2);
        rr2 = bsum(rr2);
        const double beta = rz2 / rz;
        rz = rz2; rr = rr2;
        for (int d = threadIdx.x; d < 3 * nn; d += blockDim.x) {
            const long long id = vidx(d / 3, d % 3);
            p[id] = fma(beta, p[id], dinv[id] * r[id]);
        }
        __syncthreads();
        it++;
    }
    if (threadIdx.x == 0 && its_out) *its_out = it;
}

int stencil_coarse_solve(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, const double* dinv, const double* b, double* x, double* r,
                         double* p, double* q, double rtol, int maxit) {
    if (S.fp32) k_coarse_cg<float><<<1, 1024, 0, c->stream>>>((const float*)S.A, dinv, b, x, r, p, q, g.nxn, g.nyn, g.nzn, g.px, g.nstride, g.pstride, S.nnp, rtol, maxit, S.its);
    else k_coarse_cg<double><<<1, 1024, 0, c->stream>>>((const double*)S.A, dinv, b, x, r, p, q, g.nxn, g.nyn, g.nzn, g.px, g.nstride, g.pstride, S.nnp, rtol, maxit, S.its);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}


// ---- coarsest level: explicit inverse ----------------------------------------------------------------------------------------
// The coarsest Galerkin matrix is a few hundred dofs (6x3x3 elements -> 336): it is inverted ONCE per design iteration by an
// in-place Gauss-Jordan sweep inside one CTA (SPD: no pivoting), and every V-cycle's coarse solve is one dense
// matrix-vector product in one launch -- exact, hence a linear preconditioner (the round-1 fixed-count CG was neither), and
// ~3 us instead of ~400.
template <typename T>
__global__ void __launch_bounds__(256) k_stencil_to_dense(const T* __restrict__ A, double* __restrict__ D, int nx, int ny, int nz, int px,
                                                          long long nnp, int n) {
    // one CTA per dense row (dof 3*node + c, node = i + nx (j + ny k))
    const int row = blockIdx.x, node = row / 3, c = row % 3;
    const int i = node % nx, j = (node / nx) % ny, k = node / (nx * ny);
    const long long nn = ((long long)k * ny + j) * px + i;
    for (int col = threadIdx.x; col < n; col += blockDim.x) D[(long long)row * n + col] = 0.0;
    __syncthreads();
    for (int t = threadIdx.x; t < 81; t += blockDim.x) {
        const int o = t / 3, c2 = t % 3;
        const int di = o % 3 - 1, dj = (o / 3) % 3 - 1, dk = o / 9 - 1;
        const int ii = i + di, jj = j + dj, kk = k + dk;
        if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
        D[(long long)row * n + 3 * (ii + nx * (jj + ny * kk)) + c2] = (double)A[(long long)(o * 9 + 3 * c + c2) * nnp + nn];
    }
}
// In-place Gauss-Jordan inversion (SPD, no pivoting) on G co-resident CTAs: every CTA keeps RB rows in shared memory; per pivot
// the owner publishes the scaled pivot row through a double-buffered global row, ONE software grid barrier, everybody updates its
// rows.  (A single CTA working in global memory is bound by one SM's L2 bandwidth: 5 ms for 336 dofs -- measured as a 6 ms
// regression of the multigrid leg; this version moves 2.7 KB per pivot per CTA.)
template <int RB>
__global__ void __launch_bounds__(256) k_dense_invert(double* __restrict__ D, int n, double* __restrict__ prow, unsigned int* __restrict__ cnt) {
    extern __shared__ double sm[];  // rows[RB][n]
    const int r0 = blockIdx.x * RB, G = gridDim.x;
    for (int e = threadIdx.x; e < RB * n; e += blockDim.x) {
        const int i = r0 + e / n;
        sm[e] = i < n ? D[(long long)i * n + e % n] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < n; k++) {
        double* pr = prow + (size_t)(k & 1) * n;
        if (k >= r0 && k < r0 + RB) {  // I own the pivot row: scale it, publish it
            const double* rk = sm + (k - r0) * n;
            const double p = 1.0 / rk[k];
            for (int j = threadIdx.x; j < n; j += blockDim.x) pr[j] = (j == k) ? p : rk[j] * p;
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // grid barrier k (all CTAs are resident: G <= a few per SM)
            __threadfence();
            atomicAdd(cnt, 1u);
            const long long t0 = clock64();  // (bounded: if the CTAs were ever not co-resident the inverse would be garbage, not a hang)
            while (*reinterpret_cast<volatile unsigned int*>(cnt) < (unsigned int)(k + 1) * (unsigned int)G && clock64() - t0 < 4000000000LL) {}
            __threadfence();
        }
        __syncthreads();
        const double p = __ldcg(pr + k);
        for (int e = threadIdx.x; e < RB * n; e += blockDim.x) {
            const int li = e / n, j = e - li * n, i = r0 + li;
            if (i >= n) continue;
            const double rkj = __ldcg(pr + j);
            if (i == k) sm[e] = rkj;
            else {
                const double f = sm[li * n + k];  // a_ik before this pivot's update (column k is rewritten below, after the loop barrier)
                if (j != k) sm[e] = sm[e] - f * rkj;
            }
        }
        __syncthreads();
        for (int li = threadIdx.x; li < RB; li += blockDim.x) {  // column k of my rows: -a_ik * p
            const int i = r0 + li;
            if (i < n && i != k) sm[li * n + k] = -sm[li * n + k] * p;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < RB * n; e += blockDim.x) {
        const int i = r0 + e / n;
        if (i < n) D[(long long)i * n + e % n] = sm[e];
    }
}
// x = Ainv b with b, x in the stored nodal layout; one CTA, one warp per row round-robin
__global__ void __launch_bounds__(1024) k_dense_solve(const double* __restrict__ Ainv, const double* __restrict__ b, double* __restrict__ x, int n,
                                                      int nx, int ny, int px, long long nstride, long long pstride) {
    extern __shared__ double sb[];  // b in dense dof order
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        const int node = d / 3, c = d % 3, i = node % nx, j = (node / nx) % ny, k = node / (nx * ny);
        sb[d] = b[(long long)k * pstride + c * nstride + (long long)j * px + i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int row = warp; row < n; row += nw) {
        double s = 0.0;
        for (int col = lane; col < n; col += 32) s = fma(Ainv[(long long)row * n + col], sb[col], s);
        s = warp_sum(s);
        if (lane == 0) {
            const int node = row / 3, c = row % 3, i = node % nx, j = (node / nx) % ny, k = node / (nx * ny);
            x[(long long)k * pstride + c * nstride + (long long)j * px + i] = s;
        }
    }
}

int stencil_dense_setup(tob_ctx_impl* c, const LevelGrid& g, StencilLevel& S) {
    const int n = 3 * g.nxn * g.nyn * g.nzn;
    if (n > kDenseCoarseMax) return 0;  // too large: the caller keeps the iterative coarse solve
    if (!S.Ainv) TOB_CUDA(cudaMalloc(&S.Ainv, sizeof(double) * (size_t)n * n));
    S.ndense = n;
    if (S.fp32) k_stencil_to_dense<float><<<n, 256, 0, c->stream>>>((const float*)S.A, S.Ainv, g.nxn, g.nyn, g.nzn, g.px, S.nnp, n);
    else k_stencil_to_dense<double><<<n, 256, 0, c->stream>>>((const double*)S.A, S.Ainv, g.nxn, g.nyn, g.nzn, g.px, S.nnp, n);
    if (!S.prow) {
        TOB_CUDA(cudaMalloc(&S.prow, sizeof(double) * 2 * kDenseCoarseMax));
        TOB_CUDA(cudaMalloc(&S.bar, sizeof(unsigned int)));
    }
    TOB_CUDA(cudaMemsetAsync(S.bar, 0, sizeof(unsigned int), c->stream));
    if (n <= 704) k_dense_invert<8><<<div_up(n, 8), 256, sizeof(double) * 8 * n, c->stream>>>(S.Ainv, n, S.prow, S.bar);
    else k_dense_invert<4><<<div_up(n, 4), 256, sizeof(double) * 4 * n, c->stream>>>(S.Ainv, n, S.prow, S.bar);
    c->launches += 2;
    TOB_KERNEL_CHECK();
    return 0;
}
int stencil_dense_solve(tob_ctx_impl* c, const StencilLevel& S, const LevelGrid& g, const double* b, double* x) {
    k_dense_solve<<<1, 1024, sizeof(double) * S.ndense, c->stream>>>(S.Ainv, b, x, S.ndense, g.nxn, g.nyn, g.px, g.nstride, g.pstride);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// ---- level management --------------------------------------------------------------------------------------------------------
int stencil_alloc(tob_ctx_impl* c, const LevelGrid& g, StencilLevel& S) {
    S.nnp = (long long)g.nzn * g.nyn * g.px;
    // coefficients in single precision by default: the level is a preconditioner, and its SpMV is pure coefficient streaming
    // (1944 B per node in fp64); TOB_GMG_STENCIL_FP64=1 keeps doubles (the Galerkin parity test does)
    S.fp32 = !(getenv("TOB_GMG_STENCIL_FP64") && atoi(getenv("TOB_GMG_STENCIL_FP64")) != 0);
    const size_t esz = S.fp32 ? sizeof(float) : sizeof(double);
    TOB_CUDA(cudaMalloc(&S.A, esz * 243 * S.nnp));
    TOB_CUDA(cudaMemsetAsync(S.A, 0, esz * 243 * S.nnp, c->stream));
    TOB_CUDA(cudaMalloc(&S.coef, sizeof(double) * 2 * 16));
    TOB_CUDA(cudaMalloc(&S.lam, sizeof(double)));
    TOB_CUDA(cudaMalloc(&S.its, sizeof(int)));
    TOB_CUDA(cudaMalloc(&S.pv, sizeof(double) * g.nvec));
    TOB_CUDA(cudaMalloc(&S.pt, sizeof(double) * g.nvec));
    TOB_CUDA(cudaMalloc(&S.pw, sizeof(double) * g.nvec));
    return 0;
}
void stencil_free(StencilLevel& S) {
    cudaFree(S.A); cudaFree(S.coef); cudaFree(S.lam); cudaFree(S.its); cudaFree(S.pv); cudaFree(S.pt); cudaFree(S.pw); cudaFree(S.M); cudaFree(S.Ainv); cudaFree(S.prow); cudaFree(S.bar);
    S = StencilLevel();
}

int stencil_assemble_from_mf(tob_ctx_impl* c, const LevelGrid& gf, const double* Ef, const double* KE_fine, const LevelGrid& gc,
                             const unsigned char* maskc, StencilLevel& S) {
    if (!S.M) {
        std::vector<double> M(8 * 576);
        galerkin_child_matrices(KE_fine, M.data());
        TOB_CUDA(cudaMalloc(&S.M, sizeof(double) * M.size()));
        TOB_CUDA(cudaMemcpy(S.M, M.data(), sizeof(double) * M.size(), cudaMemcpyHostToDevice));
    }
    dim3 grid(gc.nxn, gc.nyn, gc.nzn);
    if (S.fp32) k_galerkin_mf<float><<<grid, 256, 0, c->stream>>>(Ef, S.M, maskc, (float*)S.A, gc.nxn, gc.nyn, gc.nzn, gc.px, gc.pm, gc.mstride, gf.nex, gf.ney, gf.nez,
                                               gf.pe, gf.estride, S.nnp);
    else k_galerkin_mf<double><<<grid, 256, 0, c->stream>>>(Ef, S.M, maskc, (double*)S.A, gc.nxn, gc.nyn, gc.nzn, gc.px, gc.pm, gc.mstride, gf.nex, gf.ney, gf.nez,
                                               gf.pe, gf.estride, S.nnp);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}
int stencil_assemble_from_stencil(tob_ctx_impl* c, const LevelGrid& gf, const StencilLevel& F, const LevelGrid& gc, const unsigned char* maskc,
                                  StencilLevel& S) {
    dim3 grid(gc.nxn, gc.nyn, gc.nzn);
    if (F.fp32 != S.fp32) TOB_FAIL("internal: Galerkin levels with different storage precision");
    if (S.fp32) k_galerkin_st<float><<<grid, 256, 0, c->stream>>>((const float*)F.A, maskc, (float*)S.A, gf.nxn, gf.nyn, gf.nzn, gf.px, F.nnp, gc.nxn, gc.nyn, gc.nzn, gc.px, gc.pm,
                                               gc.mstride, S.nnp);
    else k_galerkin_st<double><<<grid, 256, 0, c->stream>>>((const double*)F.A, maskc, (double*)S.A, gf.nxn, gf.nyn, gf.nzn, gf.px, F.nnp, gc.nxn, gc.nyn, gc.nzn, gc.px, gc.pm,
                                               gc.mstride, S.nnp);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// dinv, lambda_max(D^-1 A) by `nit` power iterations (all on the device), Chebyshev coefficients for `deg` steps
int stencil_finish(tob_ctx_impl* c, const LevelGrid& g, StencilLevel& S, double* dinv, double ratio, int deg, int nit) {
    dim3 grid(div_up(g.nxn, 128), g.nyn, g.nzn);
    if (S.fp32) k_stencil_diag<float><<<grid, 128, 0, c->stream>>>((const float*)S.A, dinv, g.nxn, g.nyn, g.nzn, g.px, g.nstride, g.pstride, S.nnp);
    else k_stencil_diag<double><<<grid, 128, 0, c->stream>>>((const double*)S.A, dinv, g.nxn, g.nyn, g.nzn, g.px, g.nstride, g.pstride, S.nnp);
    c->launches++;
    TOB_KERNEL_CHECK();
    if (!S.pv_init) {
        // start vector: alternating signs excite the high end of the spectrum (padding stays zero).  Later calls continue from
        // the previous estimate: the design changes slowly, so the dominant eigenvector does too.
        std::vector<double> v0((size_t)g.nvec, 0.0);
        double nrm = 0.0;
        for (int k = 0; k < g.nzn; k++)
            for (int cc = 0; cc < 3; cc++)
                for (int j = 0; j < g.nyn; j++)
                    for (int i = 0; i < g.nxn; i++) {
                        const double v = (((i + j + k) & 1) ? -1.0 : 1.0) * (1.0 + 0.3 * cc) + 0.1 * ((i * 7 + j * 3 + k * 5) % 11);
                        v0[(size_t)(k * g.pstride + cc * g.nstride + (long long)j * g.px + i)] = v;
                        nrm += v * v;
                    }
        nrm = 1.0 / sqrt(nrm);
        for (auto& v : v0) v *= nrm;
        TOB_CUDA(cudaMemcpyAsync(S.pv, v0.data(), sizeof(double) * g.nvec, cudaMemcpyHostToDevice, c->stream));
        TOB_CUDA(cudaStreamSynchronize(c->stream));  // v0 goes out of scope
        S.pv_init = true;
        nit *= 3;
    }
    for (int it = 0; it < nit; it++) {
        TOB_CHECK(stencil_apply(c, S, g, EPI_Y, S.pv, S.pw, nullptr, nullptr, nullptr, 0, nullptr));
        const int nb = flat_grid(c, g.nvec);
        k_pw_scale_norm<<<nb, kFlatThreads, 0, c->stream>>>(S.pt, S.pw, dinv, g.nvec, c->d_partials);
        k_power_finish<<<1, 256, 0, c->stream>>>(S.pv, S.pt, g.nvec, c->d_partials, nb, S.lam);
        c->launches += 2;
        TOB_KERNEL_CHECK();
    }
    k_cheb_coef<<<1, 32, 0, c->stream>>>(S.lam, ratio, deg, S.coef);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

}  // namespace tob
