// gmg.cu -- geometric multigrid V-cycle preconditioner for the matrix-free PCG.
// Replaces PCMG + DMCoarsenHierarchy + DMCreateInterpolation + Galerkin PtAP + GMRES/SOR smoothers
// (LinearElasticity.cc:656-746).  B200 design:
//   * the hierarchy is the structured grid coarsened by 2 (same z-slab boundaries on every level when the
//     element layers per rank are divisible by 2^(nlvls-1)), each level is a LevelGrid, so the SAME
//     TMA z-march hex8 kernel applies the operator on every level;
//   * coarse operators are re-discretised with the 8-child average of the SIMP modulus (no assembly, no PtAP;
//     the refresh after a density change is one tiny kernel per level);
//   * smoother = Chebyshev polynomial in D^-1 A, each step ONE launch: the SpMV epilogue EPI_CHEB forms the
//     residual, the 3-term recurrence and the new iterate.  The eigenvalue bound needs no power iteration:
//     lambda_max(D^-1 A) <= lambda_max(diag(KE)^-1 KE) for ANY modulus field (element-wise Rayleigh bound);
//   * transfers are trilinear (Q1) prolongation per component and its transpose, as the reference's DMDA
//     interpolation; Dirichlet dofs are injected to the coarse levels;
//   * coarsest level: a fixed number of Jacobi-PCG iterations driven entirely by device scalars (no host sync);
//   * multi-GPU: the fine levels are z-slab distributed (halo planes over the NVLink arena); below a size threshold
//     (or where the slab boundaries stop being aligned) the levels are REPLICATED: the restricted residual of the
//     last distributed level is all-gathered once per V-cycle and every rank continues the cycle redundantly on the
//     whole coarse grid -- no halo, no all-reduce, and as many levels as the global grid allows.  Coarse levels are
//     launch-latency bound, so redundant work costs nothing while every exchange on them costs a latency.
#include "elast.cuh"
#include "galerkin.cuh"
#include "vecops.cuh"

namespace tob {

struct GmgLevel {
    LevelGrid g;
    Hex8Operator op;
    bool own_op = false;
    double *x = nullptr, *x2 = nullptr, *b = nullptr, *r = nullptr, *d = nullptr, *dinv = nullptr, *E = nullptr;
    float* dinv32 = nullptr;  // fp32 copy of dinv for the Chebyshev epilogue of the matrix-free levels (level 0: the solver's)
    unsigned char* mask = nullptr;
    bool own_E = false;
    int res = 0;  // resolution index: this level's elements are 2^res fine elements wide
    bool stencil = false;  // Galerkin level: the operator is the assembled block stencil S (galerkin.cu)
    StencilLevel S;
};

struct GmgHierarchy {
    std::vector<GmgLevel> lv;
    double lam_max = 2.0;   // bound on lambda_max(D^-1 A)
    double lam_ratio = 8.0; // Chebyshev targets [lam_max/ratio, lam_max]
    // lv[0..dI] are distributed; if dI >= 0, lv[dI+1] is the replicated twin of lv[dI] (same resolution) and the
    // levels after it are replicated too.  dI = -1: single rank, plain hierarchy.
    int dI = -1;
    std::vector<size_t> gn_off, gn_cnt, ge_off, ge_cnt, gm_off, gm_cnt;  // per-rank chunks of the gather (bytes)
    // coarsest-level PCG vectors
    double *cp = nullptr, *cq = nullptr;
    // mixed-precision smoother on the finest level (hex8_tc32.cu generation 2, tcgen05 TF32): node-interleaved fp32 vectors
    bool tf32 = false;
    float4 *xa4 = nullptr, *xb4 = nullptr, *b4 = nullptr, *d4 = nullptr, *dinv4 = nullptr;
    // CUDA graph of one V-cycle
    bool warm = false;
    bool timed = false;
    cudaGraphExec_t graph_exec = nullptr;
    const double* graph_r = nullptr;
    double* graph_z = nullptr;
    long long graph_launches = 0;
};

enum { G_S_PQ = 20, G_S_PAIR_A = 22, G_S_PAIR_B = 24 };

// ---- kernels ------------------------------------------------------------------------------------
// coarse modulus = mean of the 8 children (owned coarse layers; ghosts are exchanged afterwards)
__global__ void k_coarsen_E(const double* __restrict__ Ef, double* __restrict__ Ec, int ncx, int ncy, int pef, int pec,
                            long long esf, long long esc, int lf0, int lc0, int nlc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, l = blockIdx.z;
    if (i >= ncx || l >= nlc) return;
    double s = 0.0;
    for (int dk = 0; dk < 2; dk++)
        for (int dj = 0; dj < 2; dj++)
            for (int di = 0; di < 2; di++) s += Ef[(lf0 + 2 * l + dk) * esf + (long long)(2 * j + dj) * pef + 2 * i + di];
    Ec[(lc0 + l) * esc + (long long)j * pec + i] = 0.125 * s;
}

__global__ void k_inject_mask(const unsigned char* __restrict__ mf, unsigned char* __restrict__ mc, int ncx, int ncy, int pmf,
                              int pmc, long long msf, long long msc, int zf0, int zc0, int nzc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= ncx || k >= nzc) return;
    mc[(zc0 + k) * msc + (long long)j * pmc + i] = mf[(zf0 + 2 * k) * msf + (long long)(2 * j) * pmf + 2 * i];
}

// rc = P^T rf on owned coarse planes (fine ghost planes must be current); clamped coarse dofs get 0
__global__ void __launch_bounds__(128)
k_restrict(const double* __restrict__ rf, double* __restrict__ rc, const unsigned char* __restrict__ mc, int nfx, int nfy,
           int nfzl, int ncx, int ncy, int pxf, int pxc, int pmc, long long nsf, long long psf, long long nsc, long long psc,
           long long msc, int zc0, int nzc, int gc0, int gf0) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y, K = blockIdx.z;
    if (I >= ncx || K >= nzc) return;
    const int kc = zc0 + K;               // local coarse plane
    const int kf = 2 * (kc + gc0) - gf0;  // local fine plane coincident with it
    double acc[3] = {0, 0, 0};
    for (int dk = -1; dk <= 1; dk++) {
        const int z = kf + dk;
        if (z < 0 || z >= nfzl) continue;
        for (int dj = -1; dj <= 1; dj++) {
            const int y = 2 * J + dj;
            if (y < 0 || y >= nfy) continue;
            for (int di = -1; di <= 1; di++) {
                const int x = 2 * I + di;
                if (x < 0 || x >= nfx) continue;
                const double w = (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0);
                const long long o = z * psf + (long long)y * pxf + x;
                acc[0] = fma(w, rf[o], acc[0]);
                acc[1] = fma(w, rf[o + nsf], acc[1]);
                acc[2] = fma(w, rf[o + 2 * nsf], acc[2]);
            }
        }
    }
    const unsigned char m = mc[kc * msc + (long long)J * pmc + I];
    for (int c = 0; c < 3; c++) rc[kc * psc + c * nsc + (long long)J * pxc + I] = (m & (1 << c)) ? 0.0 : acc[c];
}

// xf += P xc on owned fine planes (coarse ghost planes must be current); clamped fine dofs stay untouched
__global__ void __launch_bounds__(128)
k_prolong_add(double* __restrict__ xf, const double* __restrict__ xc, const unsigned char* __restrict__ mf, int nfx, int nfy,
              int pxf, int pxc, int pmf, long long nsf, long long psf, long long nsc, long long psc, long long msf, int zf0,
              int nzf, int gf0, int gc0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= nfx || k >= nzf) return;
    const int kf = zf0 + k;
    const int kg = kf + gf0;  // global fine plane
    const int I0 = i >> 1, J0 = j >> 1, K0 = (kg >> 1) - gc0;
    const int ni = 1 + (i & 1), nj = 1 + (j & 1), nk = 1 + (kg & 1);
    const double w = 1.0 / (ni * nj * nk);
    double acc[3] = {0, 0, 0};
    for (int a = 0; a < nk; a++)
        for (int b = 0; b < nj; b++)
            for (int c = 0; c < ni; c++) {
                const long long o = (K0 + a) * psc + (long long)(J0 + b) * pxc + I0 + c;
                acc[0] += xc[o];
                acc[1] += xc[o + nsc];
                acc[2] += xc[o + 2 * nsc];
            }
    const unsigned char m = mf[kf * msf + (long long)j * pmf + i];
    for (int c = 0; c < 3; c++)
        if (!(m & (1 << c))) xf[kf * psf + c * nsf + (long long)j * pxf + i] += w * acc[c];
}

// first Chebyshev step from a zero iterate: d = s1 * dinv * b ; x = d
__global__ void __launch_bounds__(kFlatThreads) k_cheb_first(double* __restrict__ x, double* __restrict__ d, const double* __restrict__ b,
                                                             const double* __restrict__ dinv, double s1, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = s1 * dinv[i] * b[i];
        d[i] = v;
        x[i] = v;
    }
}

// ---- helpers ----------------------------------------------------------------------------------------
static int galloc(tob_ctx_impl* c, double** p, long long n) {
    TOB_CUDA(cudaMalloc(p, sizeof(double) * n));
    TOB_CUDA(cudaMemsetAsync(*p, 0, sizeof(double) * n, c->stream));
    return 0;
}

static double element_rayleigh_bound(const double* KE) {
    // lambda_max(diag(KE)^-1 KE) by power iteration on the 24x24 matrix (host, once)
    double v[24], w[24];
    for (int i = 0; i < 24; i++) v[i] = 1.0 + 0.37 * ((i * 7) % 5) - 0.11 * i;
    double lam = 1.0;
    for (int it = 0; it < 500; it++) {
        double nrm = 0;
        for (int i = 0; i < 24; i++) {
            double s = 0;
            for (int j = 0; j < 24; j++) s += KE[i * 24 + j] * v[j];
            w[i] = s / KE[i * 24 + i];
            nrm += w[i] * w[i];
        }
        nrm = sqrt(nrm);
        lam = nrm;
        for (int i = 0; i < 24; i++) v[i] = w[i] / nrm;
    }
    return lam;
}

static int level_apply(tob_elast_impl* h, GmgLevel& L, const SpmvArgs& a0, int* nparts, int step = 0) {
    if (L.stencil) return stencil_apply(h->c, L.S, L.g, a0.epi, a0.u, a0.y, a0.a0, a0.a1, a0.o1, step, nparts);
    SpmvArgs a = a0;
    a.E = L.E; a.mask = L.mask; a.mask_in = false;
    a.kernel = h->opts.spmv_kernel;
    if (&L == &h->mg->lv[0]) h->stats.spmv_calls++;
    return hex8_spmv(h->c, L.g, L.op, a, nparts);
}

// first Chebyshev step from a zero iterate, fused with the fp64 -> V4 conversion of the right-hand side (b4 is what the later
// steps of this V-cycle -- pre- and post-smoothing -- read)
__global__ void __launch_bounds__(256) k_cheb_first4(float4* __restrict__ x, float4* __restrict__ d, float4* __restrict__ b4,
                                                     const double* __restrict__ b, const float4* __restrict__ dinv, float s1, int nxn, int nyn,
                                                     int nzl, int px, long long nstride, long long pstride, long long p4) {
    const long long n = (long long)nxn * nyn * nzl;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % nxn);
        const long long t = i / nxn;
        const int iy = (int)(t % nyn), iz = (int)(t / nyn);
        const long long so = (long long)iz * pstride + (long long)iy * px + ix;
        const long long o = (long long)iz * p4 + (long long)iy * nxn + ix;
        const float4 bv = make_float4((float)b[so], (float)b[so + nstride], (float)b[so + 2 * nstride], 0.0f);
        const float4 dv = dinv[o];
        const float4 v = make_float4(s1 * dv.x * bv.x, s1 * dv.y * bv.y, s1 * dv.z * bv.z, 0.0f);
        b4[o] = bv;
        d[o] = v;
        x[o] = v;
    }
}

static int tf32_buffers(tob_elast_impl* h, GmgLevel& L) {
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    if (mg->xa4) return 0;
    const long long nn = v4_plane_nodes(L.g) * L.g.nzl();
    float4** vs[] = {&mg->xa4, &mg->xb4, &mg->b4, &mg->d4, &mg->dinv4};
    for (auto v : vs) {
        TOB_CUDA(cudaMalloc(v, sizeof(float4) * nn));
        TOB_CUDA(cudaMemsetAsync(*v, 0, sizeof(float4) * nn, c->stream));
    }
    return 0;
}

// The same Chebyshev-Jacobi sweep on the finest level in single precision on the tensor cores: every step is ONE launch of
// k_hex8_tc32b (operator + recurrence + new iterate).  The iterate enters / leaves in fp64 (converted), so the rest of the
// V-cycle -- residual, transfers, coarse levels -- and the outer PCG stay fp64.
static int cheb_smooth_tf32(tob_elast_impl* h, GmgLevel& L, int k, bool zero_init) {
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    const LevelGrid& g = L.g;
    const long long nn = v4_plane_nodes(g) * g.nzl();
    TOB_CHECK(tc2_prepare(h));
    // halo of a V4 vector: one plane = v4_plane_nodes float4 = 2 * v4_plane_nodes doubles (a multiple of 4) for the generic halo kernel
    LevelGrid g4 = g;
    g4.pstride = 2 * v4_plane_nodes(g);
    g4.nvec = g4.pstride * g.nzl();
    TOB_CHECK(tf32_buffers(h, L));
    const double lmax = 1.02 * mg->lam_max, lmin = mg->lam_max / mg->lam_ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    // dinv4 is refreshed with the design (gmg_update); b4 is written by the first (zero-iterate) step of the pre-smoother and is
    // still the right-hand side when the post-smoother of the same V-cycle runs
    float4 *xin = mg->xa4, *xout = mg->xb4;
    if (!zero_init) TOB_CHECK(vec_st_to_v4(c, g, L.x, xin));
    for (int s = 0; s < k; s++) {
        double s0, s1;
        if (s == 0) { s0 = 0.0; s1 = 1.0 / theta; }
        else {
            const double rho_new = 1.0 / (2.0 * sigma - rho);
            s0 = rho_new * rho; s1 = 2.0 * rho_new / delta; rho = rho_new;
        }
        if (s == 0 && zero_init) {
            k_cheb_first4<<<flat_grid(c, (long long)g.nxn * g.nyn * g.nzl()), 256, 0, c->stream>>>(
                xin, mg->d4, mg->b4, L.b, mg->dinv4, (float)s1, g.nxn, g.nyn, g.nzl(), g.px, g.nstride, g.pstride, v4_plane_nodes(g));
            c->launches++;
            TOB_KERNEL_CHECK();
            continue;
        }
        TOB_CHECK(halo_nodes(c, g4, reinterpret_cast<double*>(xin)));  // no-op on one rank
        if (getenv("TOB_TC32_GEN") && atoi(getenv("TOB_TC32_GEN")) == 2)
            TOB_CHECK(hex8_tc32b(c, g, h->tcB, xin, xout, h->E32, L.mask, 1, mg->b4, mg->dinv4, mg->d4, (float)s0, (float)s1));
        else
            TOB_CHECK(hex8_tc32c(c, g, h->tcB3, xin, xout, h->E32, L.mask, 1, mg->b4, mg->dinv4, mg->d4, (float)s0, (float)s1));
        float4* t = xin; xin = xout; xout = t;
    }
    return vec_v4_to_st(c, g, xin, L.x);
}

// k Chebyshev-Jacobi steps on A x = b.  zero_init: x is taken as 0 (saves the first SpMV).
// The iterate ends in L.x (ping-pong with L.x2).
static int cheb_smooth(tob_elast_impl* h, GmgLevel& L, int k, bool zero_init) {
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    if (mg->tf32 && &L == &mg->lv[0] && !L.stencil) return cheb_smooth_tf32(h, L, k, zero_init);
    const double lmax = 1.02 * mg->lam_max, lmin = mg->lam_max / mg->lam_ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    double* xin = L.x;
    double* xout = L.x2;
    for (int s = 0; s < k; s++) {
        double s0, s1;
        if (s == 0) { s0 = 0.0; s1 = 1.0 / theta; }
        else {
            const double rho_new = 1.0 / (2.0 * sigma - rho);
            s0 = rho_new * rho;
            s1 = 2.0 * rho_new / delta;
            rho = rho_new;
        }
        if (s == 0 && zero_init) {
            const LevelGrid& g = L.g;
            if (L.stencil) { TOB_CHECK(stencil_cheb_first(c, L.S, g, xin, L.d, L.b, L.dinv)); continue; }
            k_cheb_first<<<flat_grid(c, g.own_len), kFlatThreads, 0, c->stream>>>(xin + g.own_off, L.d + g.own_off, L.b + g.own_off,
                                                                                L.dinv + g.own_off, s1, g.own_len);
            c->launches++;
            TOB_KERNEL_CHECK();
            continue;  // iterate stays in xin
        }
        TOB_CHECK(halo_nodes(c, L.g, xin));
        SpmvArgs a;
        a.u = xin; a.y = xout; a.epi = EPI_CHEB; a.a0 = L.b; a.a1 = L.dinv; a.a1f = L.dinv32; a.o1 = L.d; a.s0 = s0; a.s1 = s1;
        TOB_CHECK(level_apply(h, L, a, nullptr, s));  // Galerkin levels read their coefficients from the device (step s)
        double* t = xin; xin = xout; xout = t;
    }
    // keep the buffer roles fixed (the V-cycle is replayed as a CUDA graph): copy back instead of swapping pointers
    if (xin != L.x) TOB_CHECK(vec_copy(c, L.x + L.g.own_off, xin + L.g.own_off, L.g.own_len));
    return 0;
}

// fixed number of Jacobi-PCG iterations on the coarsest level, x0 = 0, no host synchronisation
static int coarse_solve(tob_elast_impl* h, GmgLevel& L, int its) {
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    const LevelGrid& g = L.g;
    if (L.stencil && L.S.ndense) return stencil_dense_solve(c, L.S, g, L.b, L.x);  // explicit inverse: one launch, exact
    if (L.stencil)  // too large to invert: Jacobi-PCG to a relative residual inside one CTA (no launches, no host)
        return stencil_coarse_solve(c, L.S, g, L.dinv, L.b, L.x, L.r, mg->cp, mg->cq, 1.0e-4, 20 * its);
    TOB_CHECK(vec_zero(c, L.x, g.nvec));
    TOB_CHECK(vec_copy(c, L.r, L.b, g.nvec));
    int cur = G_S_PAIR_A, nxt = G_S_PAIR_B, np = 0;
    // slots 0/1 hold the constants 0 and 1 (elast.cu): alpha = 0 pass produces rr, rz
    TOB_CHECK(pcg_update_xr(c, g, L.x, L.r, L.r, L.r, L.dinv, 0, 1, cur));
    TOB_CHECK(pcg_update_p(c, g, mg->cp, L.r, L.dinv, 0, 1, 1));
    for (int it = 0; it < its; it++) {
        TOB_CHECK(halo_nodes(c, g, mg->cp));
        SpmvArgs a;
        a.u = mg->cp; a.y = mg->cq; a.epi = EPI_DOT;
        TOB_CHECK(level_apply(h, L, a, &np));
        TOB_CHECK(finalize_sum(c, np, 1, G_S_PQ, !g.replicated));
        TOB_CHECK(pcg_update_xr(c, g, L.x, L.r, mg->cp, mg->cq, L.dinv, cur + 1, G_S_PQ, nxt));
        TOB_CHECK(pcg_update_p(c, g, mg->cp, L.r, L.dinv, nxt + 1, cur + 1, 0));
        int t = cur; cur = nxt; nxt = t;
    }
    return 0;
}

// TOB_GMG_TIMING=1: one eager V-cycle per solve is bracketed level by level with CUDA events and printed (diagnostics for the
// latency budget of the coarse levels)
static std::vector<cudaEvent_t> g_tev;
static bool g_timing = false;
static void tmark(tob_ctx_impl* c) {
    if (!g_timing) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    g_tev.push_back(e);
}

static int vcycle(tob_elast_impl* h, int l) {
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    GmgLevel& L = mg->lv[l];
    const int nl = (int)mg->lv.size();
    tmark(c);  // entry of level l (2 marks per level: entry, and before returning)
    struct Exit { tob_ctx_impl* c; ~Exit() { tmark(c); } } exit_mark{c};
    if (l == mg->dI) {
        // last distributed level: hand the problem to its replicated twin and take back my slab of the answer
        GmgLevel& R = mg->lv[l + 1];
        TOB_CHECK(allgatherv_bytes(c, L.b + L.g.own_off, R.b, mg->gn_off.data(), mg->gn_cnt.data()));
        TOB_CHECK(vcycle(h, l + 1));
        return vec_copy(c, L.x, R.x + (long long)L.g.gn0 * L.g.pstride, L.g.nvec);  // stored planes incl. ghosts
    }
    if (l == nl - 1) return coarse_solve(h, L, h->opts.coarse_its);
    const int k = h->opts.cheb_degree;
    TOB_CHECK(cheb_smooth(h, L, k, true));
    // residual r = b - A x
    TOB_CHECK(halo_nodes(c, L.g, L.x));
    SpmvArgs a;
    a.u = L.x; a.y = L.r; a.epi = EPI_RESID; a.a0 = L.b;
    TOB_CHECK(level_apply(h, L, a, nullptr));
    // restrict
    GmgLevel& C = mg->lv[l + 1];
    TOB_CHECK(halo_nodes(c, L.g, L.r));
    {
        const LevelGrid &f = L.g, &g = C.g;
        dim3 grid(div_up(g.nxn, 128), g.nyn, g.zl_own1() - g.zl_own0());
        k_restrict<<<grid, 128, 0, c->stream>>>(L.r, C.b, C.mask, f.nxn, f.nyn, f.nzl(), g.nxn, g.nyn, f.px, g.px, g.pm, f.nstride,
                                                f.pstride, g.nstride, g.pstride, g.mstride, g.zl_own0(), g.zl_own1() - g.zl_own0(),
                                                g.gn0, f.gn0);
        c->launches++;
        TOB_KERNEL_CHECK();
    }
    TOB_CHECK(vcycle(h, l + 1));
    // prolong and correct (the gather level already holds its ghost planes)
    if (l + 1 != mg->dI) TOB_CHECK(halo_nodes(c, C.g, C.x));
    {
        const LevelGrid &f = L.g, &g = C.g;
        dim3 grid(div_up(f.nxn, 128), f.nyn, f.zl_own1() - f.zl_own0());
        k_prolong_add<<<grid, 128, 0, c->stream>>>(L.x, C.x, L.mask, f.nxn, f.nyn, f.px, g.px, f.pm, f.nstride, f.pstride, g.nstride,
                                                   g.pstride, f.mstride, f.zl_own0(), f.zl_own1() - f.zl_own0(), f.gn0, g.gn0);
        c->launches++;
        TOB_KERNEL_CHECK();
    }
    return cheb_smooth(h, L, k, false);
}

int gmg_vcycle(tob_elast_impl* h, const double* r, double* z) {
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    if (!mg) TOB_FAIL("multigrid hierarchy not set up");
    GmgLevel& L0 = mg->lv[0];
    // The V-cycle is a fixed sequence of ~100-250 small launches (coarse levels are launch-latency bound): the
    // first call runs eagerly (sets kernel attributes), the second is captured into a CUDA graph, later calls
    // replay it.  All operands (pointers, tensor maps, Chebyshev coefficients) are invariant across design
    // iterations; the graph is dropped when solver options or buffers change.
    if (mg->graph_exec && mg->graph_r == r && mg->graph_z == z) {
        TOB_CUDA(cudaGraphLaunch(mg->graph_exec, c->stream));
        c->launches += mg->graph_launches;
        return 0;
    }
    if (getenv("TOB_GMG_TIMING") && mg->warm && !mg->timed) {
        mg->timed = true;
        g_timing = true;
        g_tev.clear();
        int rc0 = vec_copy(c, L0.b, r, L0.g.nvec);
        if (!rc0) rc0 = vcycle(h, 0);
        g_timing = false;
        cudaStreamSynchronize(c->stream);
        if (c->desc.rank == 0) {
            // marks: entry(0) entry(1) ... entry(n-1) exit(n-1) exit(n-2) ... exit(0)
            const int nlv = (int)mg->lv.size();
            float tot = 0;
            cudaEventElapsedTime(&tot, g_tev.front(), g_tev.back());
            printf("[gmg timing] V-cycle %.1f us:", tot * 1e3);
            for (int l = 0; l < nlv && (int)g_tev.size() == 2 * nlv; l++) {
                float down = 0, up = 0;
                if (l + 1 < nlv) {
                    cudaEventElapsedTime(&down, g_tev[l], g_tev[l + 1]);
                    cudaEventElapsedTime(&up, g_tev[2 * nlv - 2 - l], g_tev[2 * nlv - 1 - l]);
                } else cudaEventElapsedTime(&down, g_tev[l], g_tev[l + 1]);
                printf("  L%d(%dx%dx%d%s) %.1f+%.1f", l, mg->lv[l].g.nex, mg->lv[l].g.ney, mg->lv[l].g.nez, mg->lv[l].stencil ? ",G" : "", down * 1e3, up * 1e3);
            }
            printf("\n");
            fflush(stdout);
        }
        for (auto e : g_tev) cudaEventDestroy(e);
        g_tev.clear();
        if (rc0) return rc0;
    }
    const bool capture = mg->warm && h->opts.verbose != 2;
    long long l0 = c->launches;
    if (capture) TOB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = vec_copy(c, L0.b, r, L0.g.nvec);
    if (!rc) rc = vcycle(h, 0);
    if (!rc) rc = vec_copy(c, z, L0.x, L0.g.nvec);
    if (capture) {
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) TOB_FAIL(std::string("V-cycle graph capture failed: ") + cudaGetErrorString(e));
        TOB_CUDA(cudaGraphInstantiate(&mg->graph_exec, graph, 0));
        cudaGraphDestroy(graph);
        mg->graph_r = r; mg->graph_z = z;
        mg->graph_launches = c->launches - l0;
        c->launches = l0;
        TOB_CUDA(cudaGraphLaunch(mg->graph_exec, c->stream));
        c->launches += mg->graph_launches;
        return 0;
    }
    mg->warm = true;
    return rc;
}

void gmg_drop_graph(tob_elast_impl* h) {
    GmgHierarchy* mg = h->mg;
    if (mg && mg->graph_exec) { cudaGraphExecDestroy(mg->graph_exec); mg->graph_exec = nullptr; }
}

int gmg_update(tob_elast_impl* h) {
    // refresh coarse moduli and diagonals after a density change (replaces KSPSetOperators/KSPSetUp + PtAP,
    // LinearElasticity.cc:198-200)
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    if (!mg) return 0;
    mg->lv[0].E = h->E_st;
    mg->lv[0].dinv = h->DINV;
    mg->lv[0].dinv32 = h->DINV32;
    mg->lv[0].mask = h->mask;
    if (mg->tf32) {  // the V-cycle is replayed as a CUDA graph: the fp32 copy of the modulus must be refreshed HERE, not inside it
        h->E32_valid = false;
        TOB_CHECK(tc2_prepare(h));
        TOB_CHECK(tf32_buffers(h, mg->lv[0]));
        TOB_CHECK(vec_st_to_v4(c, mg->lv[0].g, mg->lv[0].dinv, mg->dinv4));
    }
    for (size_t l = 1; l < mg->lv.size(); l++) {
        GmgLevel &F = mg->lv[l - 1], &C = mg->lv[l];
        const LevelGrid &f = F.g, &g = C.g;
        if ((int)l == mg->dI + 1) {
            // replicated twin of the last distributed level: gather the owned element layers of every rank
            TOB_CHECK(allgatherv_bytes(c, F.E + (long long)(f.ke0 - f.ge0) * f.estride, C.E, mg->ge_off.data(), mg->ge_cnt.data()));
            TOB_CHECK(hex8_diag_inv(c, g, C.op, C.E, C.mask, C.dinv, true));
            TOB_CHECK(vec_d2f(c, C.dinv, C.dinv32, g.nvec));
            continue;
        }
        if (C.stencil) {  // Galerkin product P^T A P (LinearElasticity.cc:702), redone with every new design (:198-200)
            if (F.stencil) TOB_CHECK(stencil_assemble_from_stencil(c, f, F.S, g, C.mask, C.S));
            else TOB_CHECK(stencil_assemble_from_mf(c, f, F.E, F.op.KE, g, C.mask, C.S));
            if (l + 1 == mg->lv.size()) {  // coarsest level: explicit inverse when it is small enough
                TOB_CHECK(stencil_dense_setup(c, g, C.S));
                if (C.S.ndense) continue;
            }
            TOB_CHECK(stencil_finish(c, g, C.S, C.dinv, mg->lam_ratio, std::max(1, h->opts.cheb_degree), 8));
            continue;
        }
        const int nlc = g.ke1 - g.ke0;
        dim3 grid(div_up(g.nex, 128), g.ney, nlc);
        k_coarsen_E<<<grid, 128, 0, c->stream>>>(F.E, C.E, g.nex, g.ney, f.pe, g.pe, f.estride, g.estride, f.ke0 - f.ge0,
                                                 g.ke0 - g.ge0, nlc);
        c->launches++;
        TOB_KERNEL_CHECK();
        if (!g.replicated) {
            const int own0 = g.ke0 - g.ge0, own1 = g.ke1 - g.ge0;
            TOB_CHECK(halo_elems(c, C.E, g.estride, own0, own1, 0, g.nzle() - own1, c->desc.rank > 0 ? 1 : 0, 0));
        }
        TOB_CHECK(hex8_diag_inv(c, g, C.op, C.E, C.mask, C.dinv, true));
        TOB_CHECK(vec_d2f(c, C.dinv, C.dinv32, g.nvec));
    }
    return 0;
}

// levels below this many elements are replicated on every rank instead of z-slab distributed
constexpr long long kReplicateBelowElems = 300000;

int gmg_setup(tob_elast_impl* h) {
    tob_ctx_impl* c = h->c;
    if (h->mg) return 0;
    const LevelGrid& g0 = *h->g;
    const int nres = h->opts.nlvls;
    if (nres < 2) TOB_FAIL("multigrid needs nlvls >= 2");
    const int div = 1 << (nres - 1);
    if (g0.nex % div || g0.ney % div || g0.nez % div) TOB_FAIL("MESH DIMENSION NOT COMPATIBLE WITH NUMBER OF MULTIGRID LEVELS (TopOpt.cc:182-201)");
    const int nranks = c->desc.nranks, rank = c->desc.rank;
    auto* mg = new GmgHierarchy();
    // where the hierarchy switches from distributed to replicated levels (the same decision on every rank: it only
    // depends on the ownership rule, not on communicated data)
    int dI = -1;
    std::vector<int> rk0(nranks), rk1(nranks);
    if (nranks > 1) {
        int aligned = nres - 1;  // deepest resolution index at which every slab boundary is still a coarse element boundary
        for (int r = 0; r < nranks; r++) {
            TOB_CHECK(tob_slab_partition(g0.nzn, nranks, r, nullptr, nullptr, &rk0[r], &rk1[r]));
            for (int l = 0; l <= aligned; l++)
                if ((rk0[r] % (1 << l)) || (rk1[r] % (1 << l))) { aligned = l - 1; break; }
        }
        int small = nres - 1;  // first resolution (>= 1) with fewer elements than the threshold
        for (int l = 1; l < nres; l++)
            if (((long long)(g0.nex >> l) * (g0.ney >> l) * (g0.nez >> l)) < kReplicateBelowElems) { small = l; break; }
        dI = std::min(aligned, small);
    }
    mg->dI = dI;
    // TOB_GMG_TF32=1: finest-level smoothing in single precision on the tensor cores (the interleaved fp32 iterate is halo-exchanged
    // like any nodal vector: one plane = a whole number of 32-byte units)
    mg->tf32 = getenv("TOB_GMG_TF32") && atoi(getenv("TOB_GMG_TF32")) != 0;
    const int ndist = dI >= 0 ? dI + 1 : nres;       // distributed levels (single rank: all of them)
    const int nl = dI >= 0 ? ndist + (nres - dI) : nres;
    mg->lv.resize(nl);
    mg->lam_max = element_rayleigh_bound(h->op.KE);
    GmgLevel& L0 = mg->lv[0];
    L0.g = g0;
    L0.op = h->op;
    L0.E = h->E_st; L0.dinv = h->DINV; L0.dinv32 = h->DINV32; L0.mask = h->mask;
    for (int l = 0; l < nl; l++) {
        GmgLevel& L = mg->lv[l];
        if (l > 0) {
            if (l < ndist) {  // distributed, coarsened from the level above
                const LevelGrid& f = mg->lv[l - 1].g;
                const int ke0 = f.ke0 / 2, ke1 = f.ke1 / 2;
                const int kn0 = rank == 0 ? 0 : ke0 + 1, kn1 = ke1 + 1;
                L.res = l;
                make_level_grid(L.g, f.nex / 2 + 1, f.ney / 2 + 1, f.nez / 2 + 1, 2 * f.dx, 2 * f.dy, 2 * f.dz, rank, nranks, kn0, kn1);
            } else {  // replicated: the whole grid of resolution res on every rank
                L.res = dI + (l - ndist);
                const int sc = 1 << L.res;
                const int nzn = g0.nez / sc + 1;
                make_level_grid(L.g, g0.nex / sc + 1, g0.ney / sc + 1, nzn, sc * g0.dx, sc * g0.dy, sc * g0.dz, 0, 1, 0, nzn);
                L.g.replicated = true;
            }
            TOB_CHECK(hex8_build_operator(L.g.dx, L.g.dy, L.g.dz, h->nu, L.op));
            L.own_op = true;
            TOB_CHECK(galloc(c, &L.E, L.g.evec));
            TOB_CHECK(galloc(c, &L.dinv, L.g.nvec));
            TOB_CUDA(cudaMalloc(&L.dinv32, sizeof(float) * L.g.nvec));
            TOB_CUDA(cudaMemsetAsync(L.dinv32, 0, sizeof(float) * L.g.nvec, c->stream));
            TOB_CUDA(cudaMalloc(&L.mask, (size_t)L.g.mstride * L.g.nzl()));
            TOB_CUDA(cudaMemsetAsync(L.mask, 0, (size_t)L.g.mstride * L.g.nzl(), c->stream));
            L.own_E = true;
        }
        double** vs[] = {&L.x, &L.x2, &L.b, &L.r, &L.d};
        for (auto v : vs) TOB_CHECK(galloc(c, v, L.g.nvec));
        // Galerkin levels: from resolution index gf down, where the level is small (whole grid on this rank) -- the first two
        // coarsenings stay re-discretised: with the density filter a member is several elements wide, so the 8-child mean is
        // still a faithful operator there, and an assembled level 1 would cost 4x the traffic of the matrix-free fine level
        const int gf = getenv("TOB_GMG_GALERKIN_FROM") ? atoi(getenv("TOB_GMG_GALERKIN_FROM")) : 2;
        const bool whole = (nranks == 1) || L.g.replicated;
        if (l > 0 && gf > 0 && L.res >= gf && whole && !(dI >= 0 && l == dI + 1) && (long long)L.g.nxn * L.g.nyn * L.g.nzn <= 2000000) {
            L.stencil = true;
            TOB_CHECK(stencil_alloc(c, L.g, L.S));
        }
    }
    for (int l = 1; l < nl; l++)  // a Galerkin level needs every deeper level to be Galerkin too (its children are not re-discretised from E)
        if (mg->lv[l - 1].stencil && !mg->lv[l].stencil) TOB_FAIL("internal: Galerkin levels must be contiguous down to the coarsest");
    if (dI >= 0) {  // chunks of the gather from lv[dI] (distributed) to lv[dI+1] (replicated, same pitches)
        const LevelGrid& d = mg->lv[dI].g;
        const LevelGrid& t = mg->lv[dI + 1].g;
        if (d.pstride != t.pstride || d.estride != t.estride || d.mstride != t.mstride) TOB_FAIL("internal: gather level pitches differ");
        mg->gn_off.resize(nranks); mg->gn_cnt.resize(nranks); mg->ge_off.resize(nranks); mg->ge_cnt.resize(nranks);
        mg->gm_off.resize(nranks); mg->gm_cnt.resize(nranks);
        for (int r = 0; r < nranks; r++) {
            const int ke0 = rk0[r] >> dI, ke1 = rk1[r] >> dI;
            const int kn0 = r == 0 ? 0 : ke0 + 1, kn1 = ke1 + 1;
            mg->gn_off[r] = (size_t)kn0 * d.pstride * 8; mg->gn_cnt[r] = (size_t)(kn1 - kn0) * d.pstride * 8;
            mg->ge_off[r] = (size_t)ke0 * d.estride * 8; mg->ge_cnt[r] = (size_t)(ke1 - ke0) * d.estride * 8;
            mg->gm_off[r] = (size_t)kn0 * d.mstride;     mg->gm_cnt[r] = (size_t)(kn1 - kn0) * d.mstride;
        }
    }
    TOB_CHECK(galloc(c, &mg->cp, mg->lv[nl - 1].g.nvec));
    TOB_CHECK(galloc(c, &mg->cq, mg->lv[nl - 1].g.nvec));
    h->mg = mg;
    TOB_CHECK(gmg_refresh_masks(h));
    return 0;
}

int gmg_refresh_masks(tob_elast_impl* h) {
    // Dirichlet dofs are injected level by level (all stored planes, ghosts included: the coarse plane
    // coincident with a fine ghost plane is itself a stored coarse plane)
    tob_ctx_impl* c = h->c;
    GmgHierarchy* mg = h->mg;
    if (!mg) return 0;
    mg->lv[0].mask = h->mask;
    for (size_t l = 1; l < mg->lv.size(); l++) {
        GmgLevel &F = mg->lv[l - 1], &C = mg->lv[l];
        const LevelGrid &f = F.g, &g = C.g;
        if ((int)l == mg->dI + 1) {  // replicated twin: gather the owned mask planes of every rank
            TOB_CHECK(allgatherv_bytes(c, F.mask + (long long)f.zl_own0() * f.mstride, C.mask, mg->gm_off.data(), mg->gm_cnt.data()));
            continue;
        }
        // local coarse plane kc <-> global 2*(kc+gc0) -> local fine 2*(kc+gc0)-gf0 ; handle via offset arithmetic
        for (int kc = 0; kc < g.nzl(); kc++) {
            const int kf = 2 * (kc + g.gn0) - f.gn0;
            if (kf < 0 || kf >= f.nzl()) continue;
            dim3 grid(div_up(g.nxn, 128), g.nyn, 1);
            k_inject_mask<<<grid, 128, 0, c->stream>>>(F.mask, C.mask, g.nxn, g.nyn, f.pm, g.pm, f.mstride, g.mstride, kf, kc, 1);
            c->launches++;
        }
        TOB_KERNEL_CHECK();
    }
    return 0;
}

void gmg_destroy(tob_elast_impl* h) {
    GmgHierarchy* mg = h->mg;
    if (!mg) return;
    gmg_drop_graph(h);
    for (size_t l = 0; l < mg->lv.size(); l++) {
        GmgLevel& L = mg->lv[l];
        double* vs[] = {L.x, L.x2, L.b, L.r, L.d};
        for (auto v : vs) cudaFree(v);
        if (L.own_E) { cudaFree(L.E); cudaFree(L.dinv); cudaFree(L.dinv32); cudaFree(L.mask); }
        if (L.stencil) stencil_free(L.S);
        if (L.own_op) hex8_free_operator(L.op);
    }
    cudaFree(mg->cp);
    cudaFree(mg->cq);
    cudaFree(mg->xa4); cudaFree(mg->xb4); cudaFree(mg->b4); cudaFree(mg->d4); cudaFree(mg->dinv4);
    delete mg;
    h->mg = nullptr;
}

}  // namespace tob

// diagnostics for the parity tests: the assembled block stencil of a Galerkin level, A[243][nz*ny*px] -> out; dims = {nx, ny, nz, px}
extern "C" int tob_elast_gmg_level_stencil(tob_elast* hh, int level, double* out, int* dims, int* is_stencil) {
    using namespace tob;
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    if (!h->mg) TOB_FAIL("no multigrid hierarchy (pc = 1 needed)");
    if (level < 0 || level >= (int)h->mg->lv.size()) TOB_FAIL("no such level");
    GmgLevel& L = h->mg->lv[level];
    if (dims) { dims[0] = L.g.nxn; dims[1] = L.g.nyn; dims[2] = L.g.nzn; dims[3] = L.g.px; }
    if (is_stencil) *is_stencil = L.stencil ? 1 : 0;
    if (out && L.stencil) {
        const size_t cnt = (size_t)243 * L.S.nnp;
        if (L.S.fp32) {
            std::vector<float> tmp(cnt);
            TOB_CUDA(cudaMemcpyAsync(tmp.data(), L.S.A, sizeof(float) * cnt, cudaMemcpyDeviceToHost, h->c->stream));
            TOB_CUDA(cudaStreamSynchronize(h->c->stream));
            for (size_t i = 0; i < cnt; i++) out[i] = tmp[i];
        } else {
            TOB_CUDA(cudaMemcpyAsync(out, L.S.A, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->c->stream));
            TOB_CUDA(cudaStreamSynchronize(h->c->stream));
        }
    }
    return 0;
}
extern "C" int tob_elast_gmg_nlevels(tob_elast* hh, int* n) {
    auto* h = reinterpret_cast<tob::tob_elast_impl*>(hh);
    *n = h->mg ? (int)h->mg->lv.size() : 0;
    return 0;
}
