// cgsr.cu -- the vector-update kernel of the single-reduction Jacobi-PCG (see cgsr.cuh) and peer-mapped vectors.
#include "cgsr.cuh"
#include "vecops.cuh"

namespace tob {

__device__ __forceinline__ unsigned long long cg_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// bounded spin on a peer-written epoch flag: a peer that died (or never launched) must not hang this GPU for ever -- after ~4 s
// of SM clock the wait gives up; the sums then contain garbage/NaN and the host's convergence check fails the solve loudly
__device__ __forceinline__ void cg_wait_flag(const unsigned long long* p, unsigned long long e) {
    const long long t0 = clock64();
    while (cg_ld_acquire_sys(p) < e)
        if (clock64() - t0 > 8000000000LL) break;
}
__device__ __forceinline__ void cg_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// XM: lazy update of the iterate as in vecops.cu -- 1 = even iteration of a pair (x untouched), 2 = odd iteration
// (x += alpha_even p_even + alpha p; p_even is simply the direction this kernel loads), 0 = every iteration.
// The residual is NOT stored during the iteration: u = D^-1 r carries it (r = u / dinv), so the update is
//     u -= alpha * dinv * s ;   r.u = sum u^2 / dinv ;   r.r = sum (u / dinv)^2
// and an iteration streams u, w, p, s, dinv(fp32) in and p, s, u out (+ x in/out every other iteration): 7.5 / 9.5 vector
// streams instead of 10 / 12 (device time stamps on 2 GPUs: this kernel was 52 us of an 86 us iteration).
// Set-up (MODE 0) reads the true residual r = b - A x once.
template <int MODE, int XM>
__global__ void __launch_bounds__(kFlatThreads, 4)  // 64 registers: 4 blocks per SM (the unconstrained build took 106 -> 2 blocks, 4 waves)
k_cg_update(CgDev* __restrict__ cg, int par, double* __restrict__ x, const double* __restrict__ r, double* __restrict__ u,
            const double* __restrict__ w, double* __restrict__ p, double* __restrict__ s, const float* __restrict__ dinv, long long n,
            long long plane, double* __restrict__ upart, const __grid_constant__ CgPeers pe, size_t red_off0, size_t red_off1) {
    __shared__ double red[32];
    __shared__ double sh[2];
    pdl_wait();               // everything below reads what the preceding S kernel (and the U before it) wrote
    pdl_launch_dependents();  // the S kernel's CTAs may take the SMs as this grid's blocks retire
    const CgState st = cg->st[par];
    double alpha = 0.0, beta = 0.0;
    if (MODE == 1) {
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            double gam, rr, del;
            if (blockIdx.x == 0 && lane == 0) cg_trace(pe, st.epoch + 1, 0);
            if (pe.n > 1) {
                // the three sums of every rank were stored into my arena by the last CTA of that rank's SpMV; lane k polls rank k's
                // flag (system-scope acquire loads cost ~0.6 us each: in sequence they were 5 us on 8 ranks)
                const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(pe.arena);
                double v0 = 0.0, v1 = 0.0, v2 = 0.0;
                if (lane < pe.n) {
                    cg_wait_flag(&flags[F_CG_RED0 + lane], st.epoch);
                    const size_t off = (st.epoch & 1) ? red_off1 : red_off0;
                    const double* v = reinterpret_cast<const double*>(pe.arena + off + (size_t)lane * 128 * 8) + kCgSlot;
                    v0 = __ldcg(v); v1 = __ldcg(v + 1); v2 = __ldcg(v + 2);
                }
                __syncwarp();
                gam = rr = del = 0.0;
                for (int k = 0; k < pe.n; k++) {  // rank order: bitwise identical on every rank
                    gam += __shfl_sync(0xffffffffu, v0, k); rr += __shfl_sync(0xffffffffu, v1, k); del += __shfl_sync(0xffffffffu, v2, k);
                }
            } else {
                gam = cg->sums[0]; rr = cg->sums[1]; del = cg->sums[2];
            }
            if (lane == 0) {
                if (blockIdx.x == 0) cg_trace(pe, st.epoch + 1, 1);
                double a, b;
                if (st.first != 0.0) { b = 0.0; a = gam / del; }
                else { b = gam / st.gam_prev; a = gam / (del - b * gam / st.alpha_prev); }
                sh[0] = a; sh[1] = b;
                if (blockIdx.x == 0) {
                    cg->st[par ^ 1].gam_prev = gam; cg->st[par ^ 1].alpha_prev = a; cg->st[par ^ 1].first = 0.0;
                    cg->st[par ^ 1].epoch = st.epoch + 1;
                    cg->out[0] = gam; cg->out[1] = rr; cg->out[2] = del; cg->out[3] = a;
                }
            }
        }
        __syncthreads();
        alpha = sh[0]; beta = sh[1];
    } else if (blockIdx.x == 0 && threadIdx.x == 0) {
        cg->st[par ^ 1].gam_prev = 1.0; cg->st[par ^ 1].alpha_prev = 1.0; cg->st[par ^ 1].first = 1.0;
        cg->st[par ^ 1].epoch = st.epoch + 1;
    }
    const unsigned long long e = st.epoch + 1;
    const long long n2 = n >> 1, plane2 = plane >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    const double2* r2 = reinterpret_cast<const double2*>(r);
    double2* u2 = reinterpret_cast<double2*>(u);
    const double2* w2 = reinterpret_cast<const double2*>(w);
    double2* p2 = reinterpret_cast<double2*>(p);
    double2* s2 = reinterpret_cast<double2*>(s);
    const float2* d2 = reinterpret_cast<const float2*>(dinv);
    double2* lo2 = reinterpret_cast<double2*>(pe.u_lo_ghost);
    double2* hi2 = reinterpret_cast<double2*>(pe.u_hi_ghost);
    double gam = 0.0, rr = 0.0;
    auto item = [&](const long long i) {
        const float2 df = d2[i];
        const double dx = (double)df.x, dy = (double)df.y;
        double2 un;
        if (MODE == 1) {
            const double2 uv = u2[i], wv = w2[i];
            double2 pv = p2[i], sv = s2[i];
            const double2 pe = pv;  // the previous direction (XM == 2: the even iteration's, whose x update was deferred)
            pv.x = fma(beta, pv.x, uv.x); pv.y = fma(beta, pv.y, uv.y);
            sv.x = fma(beta, sv.x, wv.x); sv.y = fma(beta, sv.y, wv.y);
            if (XM != 1) {
                double2 xv = x2[i];
                if (XM == 2) { xv.x = fma(st.alpha_prev, pe.x, xv.x); xv.y = fma(st.alpha_prev, pe.y, xv.y); }
                xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
                x2[i] = xv;
            }
            un.x = fma(-alpha * dx, sv.x, uv.x); un.y = fma(-alpha * dy, sv.y, uv.y);
            p2[i] = pv; s2[i] = sv;
        } else {
            const double2 rv = r2[i];
            un = make_double2(dx * rv.x, dy * rv.y);
        }
        u2[i] = un;
        if (lo2 && i < plane2) lo2[i] = un;                        // my first owned plane = lower neighbour's top ghost
        if (hi2 && i >= n2 - plane2) hi2[i - (n2 - plane2)] = un;  // my last owned plane = upper neighbour's bottom ghost
        // r = u / dinv (padding entries carry dinv = 0 and u = 0)
        const double rx = dx != 0.0 ? un.x / dx : 0.0, ry = dy != 0.0 ? un.y / dy : 0.0;
        gam = fma(rx, un.x, gam); gam = fma(ry, un.y, gam);
        rr = fma(rx, rx, rr); rr = fma(ry, ry, rr);
    };
    // Phase A: the two boundary planes first, by the first nbnd blocks -- their stores into the neighbours' ghost planes are
    // fenced and the neighbours' flags raised while the rest of the vector is still being streamed (the flag used to leave at
    // the very end of the kernel, behind a system fence: ~4 us per neighbour on the critical path of every iteration)
    const long long nA = min(2 * plane2, n2);
    const int nbnd = (int)min((long long)gridDim.x, (nA + blockDim.x - 1) / blockDim.x);
    if ((int)blockIdx.x < nbnd) {
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nA; j += (long long)nbnd * blockDim.x)
            item(j < plane2 ? j : n2 - nA + j);
        if (pe.n > 1 || pe.trace) {
            __syncthreads();
            if (threadIdx.x == 0) {
                fence_acq_rel_sys();  // one fence per block, behind the barrier (cumulative over the block's stores)
                const unsigned int done = atomicAdd(&cg->ucounter, 1u) + 1u;
                if (done == (unsigned int)nbnd) {
                    cg->ucounter = 0u;
                    fence_acq_rel_sys();  // tickets observed -> fence -> relaxed flag stores: one release pattern for both neighbours
                    if (pe.rank > 0) st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(pe.peer[pe.rank - 1]) + F_CG_FROM_HI, e);
                    if (pe.rank < pe.n - 1) st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(pe.peer[pe.rank + 1]) + F_CG_FROM_LO, e);
                    cg_trace(pe, e, 2);
                }
            }
        }
    }
    // Phase B: the interior, grid-stride, the boundary blocks last in the block order (they already did an item)
    const long long nI = n2 - nA;
    const int bo = ((int)blockIdx.x + (int)gridDim.x - nbnd) % (int)gridDim.x;
    for (long long k = (long long)bo * blockDim.x + threadIdx.x; k < nI; k += (long long)gridDim.x * blockDim.x) item(plane2 + k);
    gam = block_sum(gam, red);
    rr = block_sum(rr, red);
    if (threadIdx.x == 0) {
        upart[blockIdx.x] = gam;
        upart[gridDim.x + blockIdx.x] = rr;
    }
}

int cg_update(tob_ctx_impl* c, const LevelGrid& g, CgDev* cg, int par, int mode, double* x, const double* r, double* u, const double* w,
              double* p, double* s, const float* dinv, double* upart, int* nbU, const CgPeers& peers, int xmode) {
    // one grid for all variants of a solve (the S kernel's last CTA sums nb partials): a single resident wave
    int nb = resident_grid(c, (const void*)k_cg_update<1, 0>, kFlatThreads, g.own_len);
    nb = std::min(nb, resident_grid(c, (const void*)k_cg_update<1, 1>, kFlatThreads, g.own_len));
    nb = std::min(nb, resident_grid(c, (const void*)k_cg_update<1, 2>, kFlatThreads, g.own_len));
    nb = std::min(nb, resident_grid(c, (const void*)k_cg_update<0, 0>, kFlatThreads, g.own_len));
    const long long o = g.own_off;
    const size_t r0 = peers.n > 1 ? cg_red_off(peers.n, 0, 0) : 0, r1 = peers.n > 1 ? cg_red_off(peers.n, 1, 0) : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nb); cfg.blockDim = dim3(kFlatThreads); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl_enabled() && mode == 1) ? 1 : 0;
    const long long own_len = g.own_len, pstride = g.pstride;
#define TOB_CG_ARGS cg, par, x + o, r + o, u + o, w + o, p + o, s + o, dinv + o, own_len, pstride, upart, peers, r0, r1
    if (mode == 0) TOB_CUDA(cudaLaunchKernelEx(&cfg, k_cg_update<0, 0>, TOB_CG_ARGS));
    else if (xmode == 1) TOB_CUDA(cudaLaunchKernelEx(&cfg, k_cg_update<1, 1>, TOB_CG_ARGS));
    else if (xmode == 2) TOB_CUDA(cudaLaunchKernelEx(&cfg, k_cg_update<1, 2>, TOB_CG_ARGS));
    else TOB_CUDA(cudaLaunchKernelEx(&cfg, k_cg_update<1, 0>, TOB_CG_ARGS));
#undef TOB_CG_ARGS
    c->launches++;
    TOB_KERNEL_CHECK();
    if (nbU) *nbU = nb;
    return 0;
}

// ---- peer-mapped vectors ---------------------------------------------------------------------------------------------------
int peer_vec_alloc(tob_ctx_impl* c, size_t nbytes, PeerVec* out) {
    const int n = c->desc.nranks, r = c->desc.rank;
    *out = PeerVec();
    out->bytes = nbytes;
    TOB_CUDA(cudaMalloc(&out->mine, nbytes));
    TOB_CUDA(cudaMemsetAsync(out->mine, 0, nbytes, c->stream));
    out->peer[r] = out->mine;
    if (n == 1) return 0;
    cudaIpcMemHandle_t mine;
    TOB_CUDA(cudaIpcGetMemHandle(&mine, out->mine));
    unsigned char* d_all = nullptr;
    TOB_CUDA(cudaMalloc(&d_all, sizeof(mine) * n));
    TOB_CUDA(cudaMemcpyAsync(d_all + sizeof(mine) * r, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    TOB_NCCL(ncclAllGather(d_all + sizeof(mine) * r, d_all, sizeof(mine), ncclUint8, c->comm, c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<cudaIpcMemHandle_t> all(n);
    TOB_CUDA(cudaMemcpy(all.data(), d_all, sizeof(mine) * n, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    // only the z-neighbours' buffers are needed (and IPC mappings are not free): open those two
    for (int k = r - 1; k <= r + 1; k += 2) {
        if (k < 0 || k >= n) continue;
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[k], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) TOB_FAIL(std::string("cudaIpcOpenMemHandle (peer vector): ") + cudaGetErrorString(e));
        out->peer[k] = (double*)p;
    }
    return 0;
}

void peer_vec_free(tob_ctx_impl* c, PeerVec* v) {
    if (!v->mine) return;
    for (int k = 0; k < c->desc.nranks; k++)
        if (k != c->desc.rank && v->peer[k]) cudaIpcCloseMemHandle(v->peer[k]);
    cudaFree(v->mine);
    *v = PeerVec();
}

}  // namespace tob
