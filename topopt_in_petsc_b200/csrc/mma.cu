// mma.cu -- Method of Moving Asymptotes on the device, behind the reference's MMA class surface
// (MMA.h:29-140).  SURVEY 8(f) rank 1: the optimiser is the caller on the far side of the hot path
// and the Amdahl term of a design iteration once the state solve runs on the GPU.
//
// Same algorithm as MMA.cc: GenSub (asymptotes, bounds, p/q approximations, :522-649), then a dual
// interior-point Newton iteration (SolveDIP :651-688) whose O(n) passes XYZofLAMBDA / DualGrad /
// DualHess / DualResidual (:690-946) are fused into ONE streaming kernel per pass that also writes
// the primal point; the m-sized sums are reduced deterministically (per-CTA partials + finalize)
// and all-reduced over NCCL where the reference calls MPI_Allreduce (:626,764,834,925).
#include "common.cuh"
#include "vecops.cuh"

namespace tob {

constexpr int kMaxM = 8;

struct MmaPtrs {
    double* p[kMaxM];
    double* q[kMaxM];
    const double* dg[kMaxM];
};
struct MmaLam {
    double lam[kMaxM];
};

struct tob_mma_impl {
    tob_ctx_impl* c = nullptr;
    long long n = 0, nloc = 0;
    int m = 1, k = 0;
    double asyminit = 0.5, asymdec = 0.7, asyminc = 1.2;  // MMA.cc:199-201
    double a[kMaxM], cc[kMaxM], d[kMaxM];
    double y[kMaxM], z = 0, lam[kMaxM], mu[kMaxM], b[kMaxM];
    double *L = nullptr, *U = nullptr, *alpha = nullptr, *beta = nullptr, *p0 = nullptr, *q0 = nullptr;
    double *xo1 = nullptr, *xo2 = nullptr;
    MmaPtrs pq{};
    double* d_sums = nullptr;  // m + m*m device sums
    double* h_sums = nullptr;  // pinned
    DevBuf stage[6 + kMaxM];
};

__global__ void __launch_bounds__(kFlatThreads)
k_mma_movelim(const double* __restrict__ x, double* __restrict__ xmin, double* __restrict__ xmax, long long n, double Xmin,
              double Xmax, double movlim) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        xmax[i] = fmin(Xmax, v + movlim);
        xmin[i] = fmax(Xmin, v - movlim);
    }
}

// GenSub, MMA.cc:522-649.  partials[j*nblk + blk] = local part of b_j.
__global__ void __launch_bounds__(kFlatThreads)
k_mma_gensub(const double* __restrict__ x, const double* __restrict__ df, const double* __restrict__ xmin,
             const double* __restrict__ xmax, const double* __restrict__ xo1, const double* __restrict__ xo2,
             double* __restrict__ Lv, double* __restrict__ Uv, double* __restrict__ alf, double* __restrict__ bet,
             double* __restrict__ p0, double* __restrict__ q0, MmaPtrs P, int m, long long n, int kiter, double asyminit,
             double asymdec, double asyminc, double* __restrict__ partials) {
    __shared__ double red[32];
    double bsum[kMaxM];
    for (int j = 0; j < kMaxM; j++) bsum[j] = 0.0;
    const double feps = 1.0e-6;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double xv = x[i], xmi_ = xmin[i], xma_ = xmax[i];
        double L, U;
        if (kiter < 3) {  // :533-536
            L = (xv - asyminit * xma_) + asyminit * xmi_;
            U = (xv + asyminit * xma_) - asyminit * xmi_;
        } else {  // :556-573
            const double x1 = xo1[i], x2 = xo2[i];
            const double help = (xv - x1) * (x1 - x2);
            const double gamma = help < 0.0 ? asymdec : (help > 0.0 ? asyminc : 1.0);
            L = xv - gamma * (x1 - Lv[i]);
            U = xv + gamma * (Uv[i] - x1);
            const double xmi = fmax(1.0e-5, xma_ - xmi_);
            L = fmax(L, xv - 10.0 * xmi);
            L = fmin(L, xv - 0.01 * xmi);
            U = fmax(U, xv + 0.01 * xmi);
            U = fmin(U, xv + 10.0 * xmi);
        }
        Lv[i] = L;
        Uv[i] = U;
        alf[i] = fmax(xmi_, 0.9 * L + 0.1 * xv);  // :597-598
        bet[i] = fmin(xma_, 0.9 * U + 0.1 * xv);
        const double dfi = df[i];
        const double ux = U - xv, xl = xv - L, ul = U - L;
        p0[i] = ux * ux * (fmax(0.0, dfi) + 0.001 * fabs(dfi) + 0.5 * feps / ul);   // :601
        q0[i] = xl * xl * (fmax(0.0, -dfi) + 0.001 * fabs(dfi) + 0.5 * feps / ul);  // :602
        for (int j = 0; j < m; j++) {
            const double dg = P.dg[j][i];
            const double pj = ux * ux * fmax(0.0, dg);   // :612
            const double qj = xl * xl * fmax(0.0, -dg);  // :613
            P.p[j][i] = pj;
            P.q[j][i] = qj;
            bsum[j] += pj / ux + qj / xl;  // :620
        }
    }
    for (int j = 0; j < m; j++) {
        const double s = block_sum(bsum[j], red);
        if (threadIdx.x == 0) partials[(long long)j * gridDim.x + blockIdx.x] = s;
    }
}

// XYZofLAMBDA (:690-740) fused with the sums of DualGrad (:742-777) / DualResidual (:902-946) and,
// if hess != 0, DualHess (:779-880).  partials: [0,m) gradient sums, [m, m+m*m) Hessian sums.
__global__ void __launch_bounds__(kFlatThreads)
k_mma_dual_pass(double* __restrict__ x, const double* __restrict__ Lv, const double* __restrict__ Uv,
                const double* __restrict__ alf, const double* __restrict__ bet, const double* __restrict__ p0,
                const double* __restrict__ q0, MmaPtrs P, MmaLam lam, int m, long long n, int hess,
                double* __restrict__ partials) {
    __shared__ double red[32];
    double gs[kMaxM];
    double hs[kMaxM * kMaxM];
    for (int j = 0; j < kMaxM; j++) gs[j] = 0.0;
    for (int j = 0; j < kMaxM * kMaxM; j++) hs[j] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double pj = p0[i], qj = q0[i];
        double pv[kMaxM], qv[kMaxM];
        for (int j = 0; j < m; j++) {
            pv[j] = P.p[j][i];
            qv[j] = P.q[j][i];
            pj += pv[j] * lam.lam[j];
            qj += qv[j] * lam.lam[j];
        }
        const double L = Lv[i], U = Uv[i];
        const double sp = sqrt(pj), sq = sqrt(qj);
        const double xu = (sp * L + sq * U) / (sp + sq);
        double xv = xu;
        const double a = alf[i], b = bet[i];
        if (xv < a) xv = a;
        if (xv > b) xv = b;
        x[i] = xv;
        const double ux = U - xv, xl = xv - L;
        for (int j = 0; j < m; j++) gs[j] += pv[j] / ux + qv[j] / xl;
        if (hess) {
            double df2 = -1.0 / (2.0 * pj / (ux * ux * ux) + 2.0 * qj / (xl * xl * xl));
            if (xu < a || xu > b) df2 = 0.0;
            double PQ[kMaxM];
            for (int j = 0; j < m; j++) PQ[j] = pv[j] / (ux * ux) - qv[j] / (xl * xl);
            for (int j = 0; j < m; j++)
                for (int k = 0; k < m; k++) hs[j * m + k] += PQ[j] * df2 * PQ[k];
        }
    }
    for (int j = 0; j < m; j++) {
        const double s = block_sum(gs[j], red);
        if (threadIdx.x == 0) partials[(long long)j * gridDim.x + blockIdx.x] = s;
    }
    if (hess)
        for (int j = 0; j < m * m; j++) {
            const double s = block_sum(hs[j], red);
            if (threadIdx.x == 0) partials[(long long)(m + j) * gridDim.x + blockIdx.x] = s;
        }
}

__global__ void __launch_bounds__(kFlatThreads)
k_mma_design_change(const double* __restrict__ x, double* __restrict__ xold, long long n, double* __restrict__ partials) {
    __shared__ double red[32];
    double ch = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        ch = fmax(ch, fabs(v - xold[i]));
        xold[i] = v;
    }
    ch = block_max(ch, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = ch;
}

static int read_sums(tob_mma_impl* h, int n) {
    tob_ctx_impl* c = h->c;
    TOB_CUDA(cudaMemcpyAsync(h->h_sums, h->d_sums, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// lam >= 0, y, z of XYZofLAMBDA (MMA.cc:706-713)
static void host_yz(tob_mma_impl* h) {
    double lamai = 0.0;
    for (int j = 0; j < h->m; j++) {
        if (h->lam[j] < 0.0) h->lam[j] = 0.0;
        h->y[j] = fmax(0.0, h->lam[j] - h->cc[j]);
        lamai += h->lam[j] * h->a[j];
    }
    h->z = fmax(0.0, 10.0 * (lamai - 1.0));
}

static int dual_pass(tob_mma_impl* h, double* x, int hess) {
    tob_ctx_impl* c = h->c;
    host_yz(h);
    MmaLam lam;
    for (int j = 0; j < kMaxM; j++) lam.lam[j] = j < h->m ? h->lam[j] : 0.0;
    const int nb = flat_grid(c, h->nloc);
    k_mma_dual_pass<<<nb, kFlatThreads, 0, c->stream>>>(x, h->L, h->U, h->alpha, h->beta, h->p0, h->q0, h->pq, lam, h->m,
                                                      h->nloc, hess, c->d_partials);
    c->launches++;
    TOB_KERNEL_CHECK();
    const int nv = h->m + (hess ? h->m * h->m : 0);
    TOB_CHECK(finalize_sum_to(c, nb, nv, h->d_sums, true));
    return read_sums(h, nv);
}

// dense solve without pivoting (the reference's Factorize/Solve pair, MMA.cc:948-981)
static void lu_solve(double* K, double* x, int n) {
    for (int s = 0; s < n - 1; s++)
        for (int i = s + 1; i < n; i++) {
            K[i * n + s] /= K[s * n + s];
            for (int j = s + 1; j < n; j++) K[i * n + j] -= K[i * n + s] * K[s * n + j];
        }
    for (int i = 1; i < n; i++) {
        double acc = 0.0;
        for (int j = 0; j < i; j++) acc -= K[i * n + j] * x[j];
        x[i] += acc;
    }
    x[n - 1] /= K[(n - 1) * n + n - 1];
    for (int i = n - 2; i >= 0; i--) {
        double acc = x[i];
        for (int j = i + 1; j < n; j++) acc -= K[i * n + j] * x[j];
        x[i] = acc / K[i * n + i];
    }
}

// SolveDIP, MMA.cc:651-688
static int solve_dip(tob_mma_impl* h, double* x) {
    const int m = h->m;
    for (int j = 0; j < m; j++) { h->lam[j] = h->cc[j] / 2.0; h->mu[j] = 1.0; }
    const double tol = 1.0e-9 * sqrt((double)(m + h->n));
    double epsi = 1.0, err = 1.0;
    double grad[kMaxM], Hess[kMaxM * kMaxM], s[2 * kMaxM];
    while (epsi > tol) {
        int loop = 0;
        while (err > 0.9 * epsi && loop < 100) {
            loop++;
            TOB_CHECK(dual_pass(h, x, 1));
            for (int j = 0; j < m; j++) {
                grad[j] = h->h_sums[j] - h->b[j] - h->a[j] * h->z - h->y[j];  // DualGrad :768-770
                grad[j] = -1.0 * grad[j] - epsi / h->lam[j];                  // :667-669
            }
            for (int j = 0; j < m * m; j++) Hess[j] = h->h_sums[m + j];
            double lamai = 0.0;  // DualHess tail :838-866
            for (int j = 0; j < m; j++) {
                if (h->lam[j] < 0.0) h->lam[j] = 0.0;
                lamai += h->lam[j] * h->a[j];
                if (h->lam[j] > h->cc[j]) Hess[j * m + j] += -1.0;
                Hess[j * m + j] += -h->mu[j] / h->lam[j];
            }
            if (lamai > 0.0)
                for (int j = 0; j < m; j++)
                    for (int k = 0; k < m; k++) Hess[j * m + k] += -10.0 * h->a[j] * h->a[k];
            double tr = 0.0;
            for (int i = 0; i < m; i++) tr += Hess[i * m + i];
            double corr = 1e-4 * tr / m;
            if (-1.0 * corr < 1.0e-7) corr = -1.0e-7;
            for (int i = 0; i < m; i++) Hess[i * m + i] += corr;
            lu_solve(Hess, grad, m);
            for (int j = 0; j < m; j++) s[j] = grad[j];
            for (int i = 0; i < m; i++) s[m + i] = -h->mu[i] + epsi / h->lam[i] - s[i] * h->mu[i] / h->lam[i];
            double theta = 1.005;  // DualLineSearch :882-900
            for (int i = 0; i < m; i++) {
                if (theta < -1.01 * s[i] / h->lam[i]) theta = -1.01 * s[i] / h->lam[i];
                if (theta < -1.01 * s[i + m] / h->mu[i]) theta = -1.01 * s[i + m] / h->mu[i];
            }
            theta = 1.0 / theta;
            for (int i = 0; i < m; i++) {
                h->lam[i] += theta * s[i];
                h->mu[i] += theta * s[i + m];
            }
            TOB_CHECK(dual_pass(h, x, 0));
            double nrI = 0.0;  // DualResidual :930-937
            for (int j = 0; j < m; j++) {
                const double r1 = h->h_sums[j] - h->b[j] - h->a[j] * h->z - h->y[j] + h->mu[j];
                const double r2 = h->mu[j] * h->lam[j] - epsi;
                nrI = fmax(nrI, fmax(fabs(r1), fabs(r2)));
            }
            err = nrI;
        }
        epsi *= 0.1;
    }
    return 0;
}

}  // namespace tob

using namespace tob;

static int mma_in(tob_mma_impl* h, int slot, const double* p, const void** dev) {
    tob_ctx_impl* c = h->c;
    if (!p) TOB_FAIL("null vector");
    if (is_device_ptr(p)) { *dev = p; return 0; }
    const size_t nb = sizeof(double) * h->nloc;
    TOB_CHECK(h->stage[slot].reserve(nb));
    TOB_CUDA(cudaMemcpyAsync(h->stage[slot].p, p, nb, cudaMemcpyHostToDevice, c->stream));
    *dev = h->stage[slot].p;
    return 0;
}
static int mma_out_begin(tob_mma_impl* h, int slot, double* p, void** dev) {
    if (!p) TOB_FAIL("null vector");
    if (is_device_ptr(p)) { *dev = p; return 0; }
    TOB_CHECK(h->stage[slot].reserve(sizeof(double) * h->nloc));
    *dev = h->stage[slot].p;
    return 0;
}
static int mma_out_finish(tob_mma_impl* h, double* p, void* dev, bool sync) {
    if ((void*)p == dev) return 0;
    TOB_CUDA(cudaMemcpyAsync(p, dev, sizeof(double) * h->nloc, cudaMemcpyDeviceToHost, h->c->stream));
    if (sync) TOB_CUDA(cudaStreamSynchronize(h->c->stream));
    return 0;
}

extern "C" {

int tob_mma_create(tob_ctx* ctx, int m, const double* a, const double* cc, const double* d, tob_mma** out) {
    if (!ctx || !out) TOB_FAIL("null argument");
    if (m < 1 || m > kMaxM) TOB_FAIL("MMA supports 1..8 constraints");
    auto* c = reinterpret_cast<tob_ctx_impl*>(ctx);
    auto* h = new tob_mma_impl();
    h->c = c;
    h->m = m;
    h->nloc = c->g0.owned_elems();
    h->n = (long long)c->g0.nex * c->g0.ney * c->g0.nez;
    for (int j = 0; j < kMaxM; j++) {  // defaults of MMA.cc:270-274 (and TopOpt.cc:394-398)
        h->a[j] = (a && j < m) ? a[j] : 0.0;
        h->cc[j] = (cc && j < m) ? cc[j] : 1000.0;
        h->d[j] = (d && j < m) ? d[j] : 0.0;
        h->lam[j] = h->mu[j] = h->y[j] = h->b[j] = 0.0;
    }
    TOB_CUDA(cudaSetDevice(c->device));
    const size_t nb = sizeof(double) * h->nloc;
    double** vs[] = {&h->L, &h->U, &h->alpha, &h->beta, &h->p0, &h->q0, &h->xo1, &h->xo2};
    for (auto v : vs) { TOB_CUDA(cudaMalloc(v, nb)); TOB_CUDA(cudaMemsetAsync(*v, 0, nb, c->stream)); }
    for (int j = 0; j < m; j++) {
        TOB_CUDA(cudaMalloc(&h->pq.p[j], nb));
        TOB_CUDA(cudaMalloc(&h->pq.q[j], nb));
    }
    TOB_CUDA(cudaMalloc(&h->d_sums, sizeof(double) * (kMaxM + kMaxM * kMaxM)));
    TOB_CUDA(cudaMallocHost(&h->h_sums, sizeof(double) * (kMaxM + kMaxM * kMaxM)));
    *out = reinterpret_cast<tob_mma*>(h);
    return 0;
}

int tob_mma_destroy(tob_mma* hh) {
    auto* h = reinterpret_cast<tob_mma_impl*>(hh);
    if (!h) return 0;
    cudaSetDevice(h->c->device);
    cudaStreamSynchronize(h->c->stream);
    double* vs[] = {h->L, h->U, h->alpha, h->beta, h->p0, h->q0, h->xo1, h->xo2};
    for (auto v : vs) cudaFree(v);
    for (int j = 0; j < h->m; j++) { cudaFree(h->pq.p[j]); cudaFree(h->pq.q[j]); }
    cudaFree(h->d_sums);
    cudaFreeHost(h->h_sums);
    for (auto& b : h->stage) b.release();
    delete h;
    return 0;
}

int tob_mma_set_outer_movelimit(tob_mma* hh, double Xmin, double Xmax, double movlim, const double* x, double* xmin,
                                double* xmax) {
    auto* h = reinterpret_cast<tob_mma_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const void* dx; void* dmin; void* dmax;
    TOB_CHECK(mma_in(h, 0, x, &dx));
    TOB_CHECK(mma_out_begin(h, 1, xmin, &dmin));
    TOB_CHECK(mma_out_begin(h, 2, xmax, &dmax));
    k_mma_movelim<<<flat_grid(c, h->nloc), kFlatThreads, 0, c->stream>>>((const double*)dx, (double*)dmin, (double*)dmax,
                                                                       h->nloc, Xmin, Xmax, movlim);
    c->launches++;
    TOB_KERNEL_CHECK();
    TOB_CHECK(mma_out_finish(h, xmin, dmin, false));
    return mma_out_finish(h, xmax, dmax, true);
}

int tob_mma_update(tob_mma* hh, double* x, const double* dfdx, const double* gx, double** dgdx, const double* xmin,
                   const double* xmax) {
    auto* h = reinterpret_cast<tob_mma_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const int m = h->m;
    const void* ddf; const void* dmin; const void* dmax; const void* dxin;
    TOB_CHECK(mma_in(h, 0, x, &dxin));
    double* dx = (double*)dxin;  // updated in place (device copy if x is a host pointer)
    TOB_CHECK(mma_in(h, 1, dfdx, &ddf));
    TOB_CHECK(mma_in(h, 2, xmin, &dmin));
    TOB_CHECK(mma_in(h, 3, xmax, &dmax));
    for (int j = 0; j < m; j++) {
        const void* dg;
        TOB_CHECK(mma_in(h, 6 + j, dgdx[j], &dg));
        h->pq.dg[j] = (const double*)dg;
    }
    // GenSub
    h->k++;
    const int nb = flat_grid(c, h->nloc);
    k_mma_gensub<<<nb, kFlatThreads, 0, c->stream>>>(dx, (const double*)ddf, (const double*)dmin, (const double*)dmax,
                                                   h->xo1, h->xo2, h->L, h->U, h->alpha, h->beta, h->p0, h->q0, h->pq, m,
                                                   h->nloc, h->k, h->asyminit, h->asymdec, h->asyminc, c->d_partials);
    c->launches++;
    TOB_KERNEL_CHECK();
    TOB_CHECK(finalize_sum_to(c, nb, m, h->d_sums, true));
    TOB_CHECK(read_sums(h, m));
    for (int j = 0; j < m; j++) h->b[j] = h->h_sums[j] - gx[j];  // MMA.cc:630-632
    // xo2 <- xo1 <- xval (MMA.cc:512-513)
    TOB_CHECK(vec_copy(c, h->xo2, h->xo1, h->nloc));
    TOB_CHECK(vec_copy(c, h->xo1, dx, h->nloc));
    TOB_CHECK(solve_dip(h, dx));
    if ((void*)dx != (void*)x) {
        TOB_CUDA(cudaMemcpyAsync(x, dx, sizeof(double) * h->nloc, cudaMemcpyDeviceToHost, c->stream));
        TOB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int tob_mma_design_change(tob_mma* hh, const double* x, double* xold, double* ch) {
    auto* h = reinterpret_cast<tob_mma_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const void* dx; const void* dxo_in;
    TOB_CHECK(mma_in(h, 0, x, &dx));
    TOB_CHECK(mma_in(h, 1, xold, &dxo_in));
    double* dxo = (double*)dxo_in;
    const int nb = flat_grid(c, h->nloc);
    k_mma_design_change<<<nb, kFlatThreads, 0, c->stream>>>((const double*)dx, dxo, h->nloc, c->d_partials);
    c->launches++;
    TOB_KERNEL_CHECK();
    TOB_CHECK(finalize_max_to(c, nb, 1, h->d_sums, true));
    TOB_CHECK(read_sums(h, 1));
    *ch = h->h_sums[0];
    if ((void*)dxo != (void*)xold) {
        TOB_CUDA(cudaMemcpyAsync(xold, dxo, sizeof(double) * h->nloc, cudaMemcpyDeviceToHost, c->stream));
        TOB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int tob_mma_get_restart(tob_mma* hh, double* xo1, double* xo2, double* U, double* L, int* k) {
    auto* h = reinterpret_cast<tob_mma_impl*>(hh);
    const size_t nb = sizeof(double) * h->nloc;
    TOB_CUDA(cudaMemcpyAsync(xo1, h->xo1, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaMemcpyAsync(xo2, h->xo2, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaMemcpyAsync(U, h->U, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaMemcpyAsync(L, h->L, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaStreamSynchronize(h->c->stream));
    if (k) *k = h->k;
    return 0;
}
int tob_mma_set_restart(tob_mma* hh, const double* xo1, const double* xo2, const double* U, const double* L, int k) {
    auto* h = reinterpret_cast<tob_mma_impl*>(hh);
    const size_t nb = sizeof(double) * h->nloc;
    TOB_CUDA(cudaMemcpyAsync(h->xo1, xo1, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaMemcpyAsync(h->xo2, xo2, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaMemcpyAsync(h->U, U, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaMemcpyAsync(h->L, L, nb, cudaMemcpyDefault, h->c->stream));
    TOB_CUDA(cudaStreamSynchronize(h->c->stream));
    h->k = k;
    return 0;
}

}  // extern "C"
