// vecops.cu -- layout conversion at the C-ABI boundary, flat vector kernels and deterministic
// two-stage reductions.  The reference does these through PETSc Vec ops (VecAXPY, VecDot,
// VecPointwiseMult ...) inside KSPSolve (LinearElasticity.cc:204); here they are fused streaming
// kernels over the contiguous owned range of the stored layout.
#include "common.cuh"
#include "vecops.cuh"
#include <map>
#include <mutex>

namespace tob {

// ---------------------------------------------------------------------------------------
// natural <-> stored layout
// ---------------------------------------------------------------------------------------
__global__ void k_nat_to_stored_dof(const double* __restrict__ nat, double* __restrict__ st, int nxn, int nyn, int nzo,
                                    int px, long long nstride, long long pstride, int zl0) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)nxn * nyn * nzo * 3;
    if (idx >= tot) return;
    int i = idx % nxn; long long t = idx / nxn;
    int j = t % nyn; t /= nyn;
    int c = t % 3; int k = t / 3;
    st[(k + zl0) * pstride + c * nstride + (long long)j * px + i] = nat[3 * ((long long)i + (long long)nxn * (j + (long long)nyn * k)) + c];
}
__global__ void k_stored_to_nat_dof(const double* __restrict__ st, double* __restrict__ nat, int nxn, int nyn, int nzo,
                                    int px, long long nstride, long long pstride, int zl0) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)nxn * nyn * nzo * 3;
    if (idx >= tot) return;
    int i = idx % nxn; long long t = idx / nxn;
    int j = t % nyn; t /= nyn;
    int c = t % 3; int k = t / 3;
    nat[3 * ((long long)i + (long long)nxn * (j + (long long)nyn * k)) + c] = st[(k + zl0) * pstride + c * nstride + (long long)j * px + i];
}
__global__ void k_nat_to_stored_elem(const double* __restrict__ nat, double* __restrict__ st, int nex, int ney, int nzo,
                                     int pe, long long estride, int zl0) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)nex * ney * nzo;
    if (idx >= tot) return;
    int i = idx % nex; long long t = idx / nex;
    int j = t % ney; int k = t / ney;
    st[(k + zl0) * estride + (long long)j * pe + i] = nat[idx];
}
__global__ void k_stored_to_nat_elem(const double* __restrict__ st, double* __restrict__ nat, int nex, int ney, int nzo,
                                     int pe, long long estride, int zl0) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)nex * ney * nzo;
    if (idx >= tot) return;
    int i = idx % nex; long long t = idx / nex;
    int j = t % ney; int k = t / ney;
    nat[idx] = st[(k + zl0) * estride + (long long)j * pe + i];
}
__global__ void k_nat_to_stored_mask(const unsigned char* __restrict__ fixed, unsigned char* __restrict__ st, int nxn,
                                     int nyn, int nzo, int pm, int zl0) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)nxn * nyn * nzo;
    if (idx >= tot) return;
    int i = idx % nxn; long long t = idx / nxn;
    int j = t % nyn; int k = t / nyn;
    unsigned char m = 0;
    for (int c = 0; c < 3; c++) m |= (fixed[3 * idx + c] ? 1 : 0) << c;
    st[((long long)(k + zl0) * nyn + j) * pm + i] = m;
}

#define LAUNCH1D(kern, tot, ...)                                                  \
    do {                                                                          \
        long long tot__ = (tot);                                                  \
        if (tot__ > 0) {                                                          \
            kern<<<div_up(tot__, 256), 256, 0, c->stream>>>(__VA_ARGS__);         \
            c->launches++;                                                        \
            TOB_KERNEL_CHECK();                                                   \
        }                                                                         \
    } while (0)

int nat_to_stored_dof(tob_ctx_impl* c, const LevelGrid& g, const double* nat, double* st) {
    int nzo = g.kn1 - g.kn0;
    LAUNCH1D(k_nat_to_stored_dof, (long long)g.nxn * g.nyn * nzo * 3, nat, st, g.nxn, g.nyn, nzo, g.px, g.nstride,
             g.pstride, g.zl_own0());
    return 0;
}
int stored_to_nat_dof(tob_ctx_impl* c, const LevelGrid& g, const double* st, double* nat) {
    int nzo = g.kn1 - g.kn0;
    LAUNCH1D(k_stored_to_nat_dof, (long long)g.nxn * g.nyn * nzo * 3, st, nat, g.nxn, g.nyn, nzo, g.px, g.nstride,
             g.pstride, g.zl_own0());
    return 0;
}
int nat_to_stored_elem(tob_ctx_impl* c, const LevelGrid& g, const double* nat, double* st) {
    int nzo = g.ke1 - g.ke0;
    LAUNCH1D(k_nat_to_stored_elem, (long long)g.nex * g.ney * nzo, nat, st, g.nex, g.ney, nzo, g.pe, g.estride,
             g.ke0 - g.ge0);
    return 0;
}
int stored_to_nat_elem(tob_ctx_impl* c, const LevelGrid& g, const double* st, double* nat) {
    int nzo = g.ke1 - g.ke0;
    LAUNCH1D(k_stored_to_nat_elem, (long long)g.nex * g.ney * nzo, st, nat, g.nex, g.ney, nzo, g.pe, g.estride,
             g.ke0 - g.ge0);
    return 0;
}
int nat_to_stored_mask(tob_ctx_impl* c, const LevelGrid& g, const unsigned char* fixed, unsigned char* st) {
    int nzo = g.kn1 - g.kn0;
    LAUNCH1D(k_nat_to_stored_mask, (long long)g.nxn * g.nyn * nzo, fixed, st, g.nxn, g.nyn, nzo, g.pm, g.zl_own0());
    return 0;
}

// ---------------------------------------------------------------------------------------
// flat kernels
// ---------------------------------------------------------------------------------------
int vec_zero(tob_ctx_impl* c, double* v, long long n) {
    TOB_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * n, c->stream));
    return 0;
}
int vec_copy(tob_ctx_impl* c, double* dst, const double* src, long long n) {
    TOB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

__global__ void k_finalize_sum(const double* __restrict__ partials, int n, double* __restrict__ out, int nvals) {
    __shared__ double red[32];
    for (int v = 0; v < nvals; v++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[(long long)v * n + i];
        s = block_sum(s, red);
        if (threadIdx.x == 0) out[v] = s;
    }
}
__global__ void k_finalize_max(const double* __restrict__ partials, int n, double* __restrict__ out, int nvals) {
    __shared__ double red[32];
    for (int v = 0; v < nvals; v++) {
        double s = -1.0e300;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmax(s, partials[(long long)v * n + i]);
        s = block_max(s, red);
        if (threadIdx.x == 0) out[v] = s;
    }
}

int finalize_sum_to(tob_ctx_impl* c, int nblocks, int nvals, double* d_out, bool allreduce) {
    if (allreduce && c->desc.nranks > 1 && finalize_allreduce(c, nblocks, nvals, d_out, 0) == 0) return 0;
    k_finalize_sum<<<1, 256, 0, c->stream>>>(c->d_partials, nblocks, d_out, nvals);
    c->launches++;
    TOB_KERNEL_CHECK();
    if (allreduce) TOB_CHECK(allreduce_sum(c, d_out, nvals));
    return 0;
}
int finalize_max_to(tob_ctx_impl* c, int nblocks, int nvals, double* d_out, bool allreduce) {
    if (allreduce && c->desc.nranks > 1 && finalize_allreduce(c, nblocks, nvals, d_out, 1) == 0) return 0;
    k_finalize_max<<<1, 256, 0, c->stream>>>(c->d_partials, nblocks, d_out, nvals);
    c->launches++;
    TOB_KERNEL_CHECK();
    if (allreduce) TOB_CHECK(allreduce_max(c, d_out, nvals));
    return 0;
}
int finalize_sum(tob_ctx_impl* c, int nblocks, int nvals, int slot, bool allreduce) {
    return finalize_sum_to(c, nblocks, nvals, c->d_scalars + slot, allreduce);
}
int finalize_max(tob_ctx_impl* c, int nblocks, int nvals, int slot, bool allreduce) {
    return finalize_max_to(c, nblocks, nvals, c->d_scalars + slot, allreduce);
}

int flat_grid(tob_ctx_impl* c, long long n) {
    long long want = (n + 2 * kFlatThreads - 1) / (2 * kFlatThreads);
    long long cap = (long long)c->num_sms * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

int resident_grid(tob_ctx_impl* c, const void* kernel, int threads, long long n) {
    static std::mutex mu;
    static std::map<const void*, int> per_sm;
    int occ = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = per_sm.find(kernel);
        if (it == per_sm.end()) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 1;
            per_sm[kernel] = occ;
        } else occ = it->second;
    }
    long long want = (n + 2 * threads - 1) / (2 * threads);
    const long long cap = (long long)c->num_sms * occ;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

__global__ void __launch_bounds__(kFlatThreads) k_dot(const double* __restrict__ a, const double* __restrict__ b,
                                                      long long n, double* __restrict__ partials) {
    __shared__ double red[32];
    double s = 0.0;
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    long long n2 = n >> 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        double2 x = a2[i], y = b2[i];
        s = fma(x.x, y.x, s);
        s = fma(x.y, y.y, s);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) s = fma(a[n - 1], b[n - 1], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

int vec_dot(tob_ctx_impl* c, const LevelGrid& g, const double* a, const double* b, int slot) {
    int nb = flat_grid(c, g.own_len);
    k_dot<<<nb, kFlatThreads, 0, c->stream>>>(a + g.own_off, b + g.own_off, g.own_len, c->d_partials);
    c->launches++;
    TOB_KERNEL_CHECK();
    return finalize_sum(c, nb, 1, slot, true);
}

// ---- PCG fused updates ------------------------------------------------------------------
// x += alpha p ; r -= alpha q ; partial sums of r.r and (Jacobi) r.dinv.r, alpha = s[num]/s[den]
__global__ void __launch_bounds__(kFlatThreads)
k_pcg_update_xr(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                const double* __restrict__ q, const double* __restrict__ dinv, long long n,
                const double* __restrict__ scal, int s_num, int s_den, double* __restrict__ partials) {
    __shared__ double red[32];
    const double alpha = scal[s_num] / scal[s_den];
    double rr = 0.0, rz = 0.0;
    long long n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* q2 = reinterpret_cast<const double2*>(q);
    const double2* d2 = reinterpret_cast<const double2*>(dinv);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        double2 xv = x2[i], rv = r2[i], pv = p2[i], qv = q2[i];
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
        x2[i] = xv; r2[i] = rv;
        rr = fma(rv.x, rv.x, rr); rr = fma(rv.y, rv.y, rr);
        if (dinv) {
            double2 dv = d2[i];
            rz = fma(rv.x * dv.x, rv.x, rz); rz = fma(rv.y * dv.y, rv.y, rz);
        }
    }
    rr = block_sum(rr, red);
    rz = block_sum(rz, red);
    if (threadIdx.x == 0) { partials[blockIdx.x] = rr; partials[gridDim.x + blockIdx.x] = rz; }
}

// p = z + beta p with z = dinv.r (Jacobi) or an explicit z ; beta = s[num]/s[den]
__global__ void __launch_bounds__(kFlatThreads)
k_pcg_update_p(double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ dinv, long long n,
               const double* __restrict__ scal, int s_num, int s_den, int first) {
    const double beta = first ? 0.0 : scal[s_num] / scal[s_den];
    long long n2 = n >> 1;
    double2* p2 = reinterpret_cast<double2*>(p);
    const double2* r2 = reinterpret_cast<const double2*>(r);
    const double2* d2 = reinterpret_cast<const double2*>(dinv);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        double2 rv = r2[i], pv = p2[i];
        if (dinv) { double2 dv = d2[i]; rv.x *= dv.x; rv.y *= dv.y; }
        pv.x = fma(beta, pv.x, rv.x); pv.y = fma(beta, pv.y, rv.y);
        p2[i] = pv;
    }
}

int pcg_update_xr(tob_ctx_impl* c, const LevelGrid& g, double* x, double* r, const double* p, const double* q,
                  const double* dinv, int s_num, int s_den, int slot_rr_rz) {
    int nb = flat_grid(c, g.own_len);
    k_pcg_update_xr<<<nb, kFlatThreads, 0, c->stream>>>(x + g.own_off, r + g.own_off, p + g.own_off, q + g.own_off,
                                                      dinv ? dinv + g.own_off : nullptr, g.own_len, c->d_scalars,
                                                      s_num, s_den, c->d_partials);
    c->launches++;
    TOB_KERNEL_CHECK();
    return finalize_sum(c, nb, 2, slot_rr_rz, !g.replicated);
}
int pcg_update_p(tob_ctx_impl* c, const LevelGrid& g, double* p, const double* z_or_r, const double* dinv, int s_num,
                 int s_den, int first) {
    int nb = flat_grid(c, g.own_len);
    k_pcg_update_p<<<nb, kFlatThreads, 0, c->stream>>>(p + g.own_off, z_or_r + g.own_off,
                                                     dinv ? dinv + g.own_off : nullptr, g.own_len, c->d_scalars, s_num,
                                                     s_den, first);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// y = a*x + b*y  /  w = a*x + b*y  style helpers (host scalars)
__global__ void __launch_bounds__(kFlatThreads) k_axpby(double* __restrict__ y, const double* __restrict__ x, double a,
                                                        double b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = a * x[i] + b * y[i];
}
int vec_axpby(tob_ctx_impl* c, double* y, const double* x, double a, double b, long long n) {
    if (n <= 0) return 0;
    k_axpby<<<flat_grid(c, n), kFlatThreads, 0, c->stream>>>(y, x, a, b, n);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}
__global__ void __launch_bounds__(kFlatThreads) k_pointwise_mult(double* __restrict__ y, const double* __restrict__ a,
                                                                 const double* __restrict__ b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = a[i] * b[i];
}
int vec_pointwise_mult(tob_ctx_impl* c, double* y, const double* a, const double* b, long long n) {
    if (n <= 0) return 0;
    k_pointwise_mult<<<flat_grid(c, n), kFlatThreads, 0, c->stream>>>(y, a, b, n);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}
// y = free ? y : u   (restores the (I-N) u part of the Dirichlet operator on clamped dofs)
__global__ void __launch_bounds__(kFlatThreads) k_restore_fixed(double* __restrict__ y, const double* __restrict__ u,
                                                                const double* __restrict__ nf, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (nf[i] == 0.0) y[i] = u[i];
}
int vec_restore_fixed(tob_ctx_impl* c, double* y, const double* u, const double* nf, long long n) {
    if (n <= 0) return 0;
    k_restore_fixed<<<flat_grid(c, n), kFlatThreads, 0, c->stream>>>(y, u, nf, n);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}
__global__ void __launch_bounds__(kFlatThreads) k_fill(double* __restrict__ y, double v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = v;
}
int vec_fill(tob_ctx_impl* c, double* y, double v, long long n) {
    if (n <= 0) return 0;
    k_fill<<<flat_grid(c, n), kFlatThreads, 0, c->stream>>>(y, v, n);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// ---- single-rank Jacobi-PCG without finalize kernels --------------------------------------------------------
// The two 1-CTA finalize launches per iteration (and their dependency gaps) cost ~18 us of a ~157 us iteration at
// 192x96x96 (profiles/r01_bench_launches_cfg2.csv).  On one GPU the consumers can reduce the producer's per-CTA
// partial sums themselves: every CTA sums the same <= few-thousand doubles in the same order, so all CTAs compute
// bitwise identical alpha / beta.  Layout of an "xr partial block": [0,nb) = r.r, [nb,2nb) = r.dinv.r.
__device__ __forceinline__ double sum_parts_bcast(const double* __restrict__ p, int n, double* red, double* bc) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += p[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *bc = s;
    __syncthreads();
    s = *bc;
    __syncthreads();
    return s;
}

// H = L2 hint mask (common.cuh): bit 0 streams evict_first, bit 1 p evict_last, bit 2 q last-use / r keep
// XM = how the iterate is updated.  0: x += alpha p every iteration.  1 / 2: LAZY x -- iterations come in (even, odd) pairs; the
// even one leaves x alone (it does not even read x or p: 3.5 instead of 6.5 vector streams) and parks its alpha in ab[0]; the odd
// one applies both updates at once,  x += alpha_even p_even + alpha p,  recovering  p_even = (p - z) / beta  from the direction
// recurrence p = z + beta p_even (z = D^-1 r with the r this kernel starts from, beta = ab[1] left by the p-update kernel).
// A pair moves 10 vector streams instead of 13.  The solver only stops after complete pairs.
template <int H, int XM>
__global__ void __launch_bounds__(kFlatThreads)
k_pcg_update_xr_f(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ q,
                  const float* __restrict__ dinv, long long n, const double* __restrict__ pq_parts, int n_pq,
                  const double* __restrict__ prev_parts, int nb_prev, double* __restrict__ out_parts, double* __restrict__ ab) {
    __shared__ double red[32];
    __shared__ double bc;
    double alpha = 0.0;  // pq_parts == nullptr: the initial pass that only forms r.r and r.dinv.r
    if (pq_parts) {
        const double pq = sum_parts_bcast(pq_parts, n_pq, red, &bc);
        const double rz = sum_parts_bcast(prev_parts + nb_prev, nb_prev, red, &bc);
        alpha = rz / pq;
    }
    double a_even = 0.0, inv_beta = 0.0;
    if (XM == 1 && blockIdx.x == 0 && threadIdx.x == 0) ab[0] = alpha;
    if (XM == 2) { a_even = ab[0]; inv_beta = 1.0 / ab[1]; }
    const uint64_t pf = l2_policy_first(), pl = l2_policy_last();
    double rr = 0.0, rz2 = 0.0;
    long long n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* q2 = reinterpret_cast<const double2*>(q);
    const float2* d2 = reinterpret_cast<const float2*>(dinv);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        double2 rv = (H & 1) ? ld2_hint(r2 + i, pf) : r2[i];
        const float2 dvf = __ldcs(d2 + i);  // the Jacobi preconditioner is stored in single precision (any SPD diagonal is a valid
        const double2 dv = make_double2((double)dvf.x, (double)dvf.y);  // preconditioner; 4 instead of 8 bytes per dof in two of the three kernels)
        if (pq_parts) {
            const double2 qv = (H & 4) ? ld2_hint(q2 + i, pf) : q2[i];
            if (XM != 1) {
                double2 xv = (H & 1) ? ld2_hint(x2 + i, pf) : x2[i];
                const double2 pv = (H & 2) ? ld2_hint(p2 + i, pl) : p2[i];
                if (XM == 2) {  // the deferred update of the even iteration: p_even = (p - D^-1 r) / beta
                    xv.x = fma(a_even * inv_beta, pv.x - dv.x * rv.x, xv.x);
                    xv.y = fma(a_even * inv_beta, pv.y - dv.y * rv.y, xv.y);
                }
                xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
                if (H & 1) st2_hint(x2 + i, xv, pf); else x2[i] = xv;
            }
            rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
            if ((H & 1) && !(H & 4)) st2_hint(r2 + i, rv, pf); else r2[i] = rv;  // bit 2: keep the new r for update_p
        }
        rr = fma(rv.x, rv.x, rr); rr = fma(rv.y, rv.y, rr);
        rz2 = fma(rv.x * dv.x, rv.x, rz2); rz2 = fma(rv.y * dv.y, rv.y, rz2);
    }
    rr = block_sum(rr, red);
    rz2 = block_sum(rz2, red);
    if (threadIdx.x == 0) { out_parts[blockIdx.x] = rr; out_parts[gridDim.x + blockIdx.x] = rz2; }
}

template <int H>
__global__ void __launch_bounds__(kFlatThreads)
k_pcg_update_p_f(double* __restrict__ p, const double* __restrict__ r, const float* __restrict__ dinv, long long n,
                 const double* __restrict__ new_parts, const double* __restrict__ old_parts, int nb, double* __restrict__ ab) {
    __shared__ double red[32];
    __shared__ double bc;
    double beta = 0.0;  // old_parts == nullptr: first direction p = dinv r
    if (old_parts) {
        const double rz_new = sum_parts_bcast(new_parts + nb, nb, red, &bc);
        const double rz_old = sum_parts_bcast(old_parts + nb, nb, red, &bc);
        beta = rz_new / rz_old;
        if (ab && blockIdx.x == 0 && threadIdx.x == 0) ab[1] = beta;  // the lazy-x update of the next iteration needs it
    }
    const uint64_t pf = l2_policy_first(), pl = l2_policy_last();
    long long n2 = n >> 1;
    double2* p2 = reinterpret_cast<double2*>(p);
    const double2* r2 = reinterpret_cast<const double2*>(r);
    const float2* d2 = reinterpret_cast<const float2*>(dinv);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const double2 rv = (H & 5) ? ld2_hint(r2 + i, pf) : r2[i];  // last use of r in this iteration
        const float2 dvf = __ldcs(d2 + i);
        const double2 dv = make_double2((double)dvf.x, (double)dvf.y);
        double2 pv = old_parts ? ((H & 2) ? ld2_hint(p2 + i, pl) : p2[i]) : make_double2(0.0, 0.0);
        pv.x = fma(beta, pv.x, rv.x * dv.x); pv.y = fma(beta, pv.y, rv.y * dv.y);
        if (H & 2) st2_hint(p2 + i, pv, pl); else p2[i] = pv;
    }
}

#define TOB_L2HINT_SWITCH(mask, CALL) \
    switch ((mask) & 7) {             \
        case 1: { constexpr int HM = 1; CALL; } break; \
        case 3: { constexpr int HM = 3; CALL; } break; \
        case 7: { constexpr int HM = 7; CALL; } break; \
        default: { constexpr int HM = 0; CALL; } break; \
    }

int pcg_update_xr_f(tob_ctx_impl* c, const LevelGrid& g, double* x, double* r, const double* p, const double* q, const float* dinv,
                    const double* pq_parts, int n_pq, const double* prev_parts, double* out_parts, int* nb_out, int xmode, double* ab) {
    // every launch of the solve uses the same grid (the partial-sum blocks of one launch are read by the next): the occupancy
    // of the heaviest variant decides
    int nb = 1 << 30;
    TOB_L2HINT_SWITCH(c->l2hint, (nb = std::min(nb, resident_grid(c, (const void*)k_pcg_update_xr_f<HM, 0>, kFlatThreads, g.own_len)),
                                  nb = std::min(nb, resident_grid(c, (const void*)k_pcg_update_xr_f<HM, 1>, kFlatThreads, g.own_len)),
                                  nb = std::min(nb, resident_grid(c, (const void*)k_pcg_update_xr_f<HM, 2>, kFlatThreads, g.own_len))));
#define TOB_XR_ARGS x + g.own_off, r + g.own_off, p + g.own_off, q + g.own_off, dinv + g.own_off, g.own_len, pq_parts, n_pq, prev_parts, nb, out_parts, ab
    if (xmode == 1) { TOB_L2HINT_SWITCH(c->l2hint, (k_pcg_update_xr_f<HM, 1><<<nb, kFlatThreads, 0, c->stream>>>(TOB_XR_ARGS))); }
    else if (xmode == 2) { TOB_L2HINT_SWITCH(c->l2hint, (k_pcg_update_xr_f<HM, 2><<<nb, kFlatThreads, 0, c->stream>>>(TOB_XR_ARGS))); }
    else { TOB_L2HINT_SWITCH(c->l2hint, (k_pcg_update_xr_f<HM, 0><<<nb, kFlatThreads, 0, c->stream>>>(TOB_XR_ARGS))); }
#undef TOB_XR_ARGS
    c->launches++;
    TOB_KERNEL_CHECK();
    if (nb_out) *nb_out = nb;
    return 0;
}
int pcg_update_p_f(tob_ctx_impl* c, const LevelGrid& g, double* p, const double* r, const float* dinv, const double* new_parts,
                   const double* old_parts, double* ab, int nparts) {
    int nb = 1;
    TOB_L2HINT_SWITCH(c->l2hint, (nb = resident_grid(c, (const void*)k_pcg_update_p_f<HM>, kFlatThreads, g.own_len)));
    TOB_L2HINT_SWITCH(c->l2hint, (k_pcg_update_p_f<HM><<<nb, kFlatThreads, 0, c->stream>>>(p + g.own_off, r + g.own_off, dinv + g.own_off,
                                                                                     g.own_len, new_parts, old_parts, nparts, ab)));
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}
// sum an xr partial block into two consecutive scalar slots (host convergence checks)
int finalize_parts(tob_ctx_impl* c, const double* parts, int nb, int slot) {
    k_finalize_sum<<<1, 256, 0, c->stream>>>(parts, nb, c->d_scalars + slot, 2);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

}  // namespace tob

extern "C" int tob_vec_scale(tob_ctx* ctx, double* v, double a, long long n) {
    using namespace tob;
    auto* c = reinterpret_cast<tob_ctx_impl*>(ctx);
    const void* din;
    TOB_CHECK(stage_in(c, 0, v, sizeof(double) * n, &din));
    double* d = (double*)din;
    TOB_CHECK(vec_axpby(c, d, d, a, 0.0, n));
    if ((void*)d != (void*)v) {
        TOB_CUDA(cudaMemcpyAsync(v, d, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        TOB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}
