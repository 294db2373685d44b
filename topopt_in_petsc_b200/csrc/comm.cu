// comm.cu -- z-slab halo exchange and scalar all-reduce.
// Replaces DMGlobalToLocalBegin/End (LinearElasticity.cc:249-250,317-318,389-390), the VecScatter inside
// MatMult, and the MPI_Allreduce calls of LinearElasticity.cc:283,429 / Filter.cc:221 / MMA.cc:626,764,834,925.
//
// Two transports:
//  * NVLink peer memory (default for the hot collectives): every rank exports one cudaMalloc'ed arena through
//    CUDA IPC; a halo exchange is ONE kernel per rank that stores its boundary planes straight into the
//    neighbours' arenas over NVLink, publishes a system-scope epoch flag, waits for the neighbours' flags and
//    unpacks its own ghost planes; a scalar all-reduce is fused into the reduction's finalize kernel (every rank
//    stores its partial into every peer's arena, then sums the n contributions in rank order -> bitwise identical
//    results on all ranks).  Epoch counters live in device memory, so the kernels can be captured in CUDA graphs.
//    Measured motivation: with NCCL send/recv + 8-byte ncclAllReduce per PCG iteration the 192x96x96 job ran
//    3.3 / 4.2 / 5.3 / 4.7 design it/s on 1/2/4/8 GPUs (profiles/r01_bench_n*_v2.json): pure launch/latency.
//  * NCCL (set-up traffic, wide element-field halos of the cone filter, and everything when TOB_P2P=0).
#include "common.cuh"
#include "vecops.cuh"
#include "cgsr.cuh"

namespace tob {

// ---- arena layout -------------------------------------------------------------------------------------
constexpr size_t kFlagBytes = 4096;                       // u64 counters and flags
constexpr size_t kRedSlots = 128;                         // doubles per rank per parity
constexpr int kHaloBlocks = 64, kHaloThreads = 256;       // always fully resident (the kernel spins on remote flags)
// u64 indices inside the flag page
enum { F_HALO_EPOCH = 0, F_HALO_CTR1 = 1, F_HALO_CTR2 = 2, F_RED_EPOCH = 3, F_HALO_FROM_LO = 8, F_HALO_FROM_HI = 9, F_RED_FLAG0 = 16 };

struct PeerPtrs {
    unsigned char* p[kMaxRanks];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ size_t red_off(int n, int parity, int rank) { return kFlagBytes + ((size_t)(parity * n + rank) * kRedSlots) * 8; }
size_t cg_red_off(int n, int parity, int rank) { return kFlagBytes + ((size_t)(parity * n + rank) * kRedSlots) * 8; }  // == red_off (cgsr.cuh)
static size_t zone_off(int n, size_t plane_bytes, int parity, int dir) {
    return kFlagBytes + (size_t)2 * n * kRedSlots * 8 + (size_t)(parity * 2 + dir) * plane_bytes;
}

// One launch per rank: push my boundary planes into the neighbours' arenas, signal, wait, unpack my ghosts.
// lo/hi = arenas of the lower/upper neighbour (null if none); zone dir 0 = "data coming from my lower neighbour".
__global__ void __launch_bounds__(kHaloThreads)
k_halo_p2p(double* __restrict__ v, long long plane, int zl_first_owned, int zl_last_owned, int zl_ghost_lo, int zl_ghost_hi,
           unsigned char* me, unsigned char* lo, unsigned char* hi, size_t zone_lo_in_peer0, size_t zone_hi_in_peer0,
           size_t zone_from_lo0, size_t zone_from_hi0, size_t plane_bytes_alloc) {
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(me);
    const unsigned long long e = flags[F_HALO_EPOCH] + 1;  // incremented by the last block when everybody is done
    const size_t par = (size_t)(e & 1) * 2 * plane_bytes_alloc;
    const long long n2 = plane >> 1;  // planes are multiples of 4 doubles
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // phase 1: NVLink stores into the neighbours
    if (lo) {  // my first owned plane is the lower neighbour's upper ghost (its "from hi" zone)
        const double2* src = reinterpret_cast<const double2*>(v + (long long)zl_first_owned * plane);
        double2* dst = reinterpret_cast<double2*>(lo + zone_hi_in_peer0 + par);
        for (long long i = t0; i < n2; i += stride) dst[i] = src[i];
    }
    if (hi) {
        const double2* src = reinterpret_cast<const double2*>(v + (long long)zl_last_owned * plane);
        double2* dst = reinterpret_cast<double2*>(hi + zone_lo_in_peer0 + par);
        for (long long i = t0; i < n2; i += stride) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(&flags[F_HALO_CTR1], 1ull) + 1;
        if (done == e * gridDim.x) {  // every block of this launch has pushed its share
            __threadfence_system();
            if (lo) st_release_sys(reinterpret_cast<unsigned long long*>(lo) + F_HALO_FROM_HI, e);
            if (hi) st_release_sys(reinterpret_cast<unsigned long long*>(hi) + F_HALO_FROM_LO, e);
        }
        // phase 2: wait for the neighbours' data of this epoch
        const long long t0 = clock64();  // bounded spins (~4 s): a dead peer must not hang this GPU for ever
        if (lo) while (ld_acquire_sys(&flags[F_HALO_FROM_LO]) < e && clock64() - t0 < 8000000000LL) {}
        if (hi) while (ld_acquire_sys(&flags[F_HALO_FROM_HI]) < e && clock64() - t0 < 8000000000LL) {}
    }
    __syncthreads();
    if (lo) {
        const double2* src = reinterpret_cast<const double2*>(me + zone_from_lo0 + par);
        double2* dst = reinterpret_cast<double2*>(v + (long long)zl_ghost_lo * plane);
        for (long long i = t0; i < n2; i += stride) dst[i] = __ldcg(src + i);  // written by the peer: read through L2
    }
    if (hi) {
        const double2* src = reinterpret_cast<const double2*>(me + zone_from_hi0 + par);
        double2* dst = reinterpret_cast<double2*>(v + (long long)zl_ghost_hi * plane);
        for (long long i = t0; i < n2; i += stride) dst[i] = __ldcg(src + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(&flags[F_HALO_CTR2], 1ull) + 1;
        if (done == e * gridDim.x) flags[F_HALO_EPOCH] = e;  // all blocks have read the old epoch by now
    }
}

// finalize + all-reduce in one launch (1 block): reduce the per-CTA partials, exchange through peer stores,
// combine in rank order.  op 0 = sum, 1 = max.
__global__ void __launch_bounds__(256)
k_finalize_p2p(const double* __restrict__ partials, int nblk, double* __restrict__ out, int nvals, int op, int n, int rank,
               unsigned char* me, PeerPtrs peers) {
    __shared__ double red[32];
    __shared__ double vals[kRedSlots];
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(me);
    const unsigned long long e = flags[F_RED_EPOCH] + 1;
    const int par = (int)(e & 1);
    for (int v = 0; v < nvals; v++) {
        double s = op ? -1.0e300 : 0.0;
        if (partials) {
            for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
                const double p = partials[(long long)v * nblk + i];
                s = op ? fmax(s, p) : s + p;
            }
            s = op ? block_max(s, red) : block_sum(s, red);
        } else if (threadIdx.x == 0) s = out[v];  // in-place all-reduce of values already in `out`
        if (threadIdx.x == 0) vals[v] = s;
    }
    __syncthreads();
    // my contribution into slot [par][rank] of every arena (my own included)
    for (int idx = threadIdx.x; idx < n * nvals; idx += blockDim.x) {
        const int r = idx / nvals, i = idx % nvals;
        reinterpret_cast<double*>(peers.p[r] + red_off(n, par, rank))[i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < n) {
        st_release_sys(reinterpret_cast<unsigned long long*>(peers.p[threadIdx.x]) + F_RED_FLAG0 + rank, e);
        const long long t0 = clock64();
        while (ld_acquire_sys(&flags[F_RED_FLAG0 + threadIdx.x]) < e && clock64() - t0 < 8000000000LL) {}
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nvals; i += blockDim.x) {
        double s = __ldcg(reinterpret_cast<const double*>(me + red_off(n, par, 0)) + i);
        for (int r = 1; r < n; r++) {
            const double p = __ldcg(reinterpret_cast<const double*>(me + red_off(n, par, r)) + i);
            s = op ? fmax(s, p) : s + p;
        }
        out[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) flags[F_RED_EPOCH] = e;
}

// ---- set-up ----------------------------------------------------------------------------------------------
int peer_setup(tob_ctx_impl* c, size_t plane_bytes) {
    const int n = c->desc.nranks, r = c->desc.rank;
    if (n == 1) return 0;
    const char* env = getenv("TOB_P2P");
    if (env && atoi(env) == 0) return 0;
    if (n > kMaxRanks) TOB_FAIL("too many ranks for the peer arena");
    PeerComm& pc = c->peer;
    pc.plane_bytes = (plane_bytes + 255) / 256 * 256;
    pc.bytes = kFlagBytes + (size_t)2 * n * kRedSlots * 8 + 4 * pc.plane_bytes;
    TOB_CUDA(cudaMalloc(&pc.arena, pc.bytes));
    TOB_CUDA(cudaMemset(pc.arena, 0, pc.bytes));
    // exchange the IPC handles with an NCCL all-gather of raw bytes
    cudaIpcMemHandle_t mine;
    TOB_CUDA(cudaIpcGetMemHandle(&mine, pc.arena));
    unsigned char* d_all = nullptr;
    TOB_CUDA(cudaMalloc(&d_all, sizeof(mine) * n));
    TOB_CUDA(cudaMemcpy(d_all + sizeof(mine) * r, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    TOB_NCCL(ncclAllGather(d_all + sizeof(mine) * r, d_all, sizeof(mine), ncclUint8, c->comm, c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<cudaIpcMemHandle_t> all(n);
    TOB_CUDA(cudaMemcpy(all.data(), d_all, sizeof(mine) * n, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    // Open every peer's arena; the decision to use the peer transport is COLLECTIVE (a rank that cannot map a peer must not
    // leave the others spinning on flags it will never write): all-reduce(min) of the local outcome over NCCL.
    int ok_local = 1;
    cudaError_t first_err = cudaSuccess;
    for (int k = 0; k < n; k++) {
        if (k == r) { pc.peer[k] = pc.arena; continue; }
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[k], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (ok_local) first_err = e;
            ok_local = 0;
            pc.peer[k] = nullptr;
            continue;
        }
        pc.peer[k] = (unsigned char*)p;
    }
    {
        int* d_ok = nullptr;
        TOB_CUDA(cudaMalloc(&d_ok, sizeof(int)));
        TOB_CUDA(cudaMemcpy(d_ok, &ok_local, sizeof(int), cudaMemcpyHostToDevice));
        TOB_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c->comm, c->stream));
        TOB_CUDA(cudaStreamSynchronize(c->stream));
        int ok_all = 0;
        TOB_CUDA(cudaMemcpy(&ok_all, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
        cudaFree(d_ok);
        if (!ok_all) {
            if (!ok_local || r == 0)
                fprintf(stderr, "libtopopt_b200[rank %d]: CUDA IPC peer access unavailable on %s (%s); every rank uses NCCL for halos\n", r,
                        ok_local ? "another rank" : "this rank", ok_local ? "-" : cudaGetErrorString(first_err));
            for (int j = 0; j < n; j++) if (j != r && pc.peer[j]) cudaIpcCloseMemHandle(pc.peer[j]);
            cudaFree(pc.arena);
            pc = PeerComm();
            return 0;
        }
    }
    pc.enabled = true;
    // nobody may touch a peer arena before every rank has zeroed its own
    double* d_dummy = c->d_scalars + kScalarSlots - 1;
    TOB_NCCL(ncclAllReduce(d_dummy, d_dummy, 1, ncclDouble, ncclSum, c->comm, c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

void peer_teardown(tob_ctx_impl* c) {
    PeerComm& pc = c->peer;
    if (!pc.arena) return;
    for (int k = 0; k < c->desc.nranks; k++)
        if (k != c->desc.rank && pc.peer[k]) cudaIpcCloseMemHandle(pc.peer[k]);
    cudaFree(pc.arena);
    pc = PeerComm();
}

// ---- halo exchange of nodal planes -------------------------------------------------------------------------
int halo_nodes(tob_ctx_impl* c, const LevelGrid& g, double* v) {
    if (c->desc.nranks == 1 || g.replicated) return 0;
    const int r = c->desc.rank, n = c->desc.nranks;
    const size_t cnt = (size_t)g.pstride;
    PeerComm& pc = c->peer;
    if (pc.enabled && cnt * 8 <= pc.plane_bytes && (cnt % 4) == 0) {
        unsigned char* lo = r > 0 ? pc.peer[r - 1] : nullptr;
        unsigned char* hi = r < n - 1 ? pc.peer[r + 1] : nullptr;
        k_halo_p2p<<<kHaloBlocks, kHaloThreads, 0, c->stream>>>(v, g.pstride, g.zl_own0(), g.zl_own1() - 1, 0, g.nzl() - 1, pc.arena, lo, hi,
                                                                zone_off(n, pc.plane_bytes, 0, 0), zone_off(n, pc.plane_bytes, 0, 1),
                                                                zone_off(n, pc.plane_bytes, 0, 0), zone_off(n, pc.plane_bytes, 0, 1),
                                                                pc.plane_bytes);
        c->launches++;
        TOB_KERNEL_CHECK();
        return 0;
    }
    TOB_NCCL(ncclGroupStart());
    if (r > 0) {  // lower neighbour: send my first owned plane, receive its last owned plane into local plane 0
        TOB_NCCL(ncclSend(v + (long long)g.zl_own0() * g.pstride, cnt, ncclDouble, r - 1, c->comm, c->stream));
        TOB_NCCL(ncclRecv(v, cnt, ncclDouble, r - 1, c->comm, c->stream));
    }
    if (r < n - 1) {  // upper neighbour
        TOB_NCCL(ncclSend(v + (long long)(g.zl_own1() - 1) * g.pstride, cnt, ncclDouble, r + 1, c->comm, c->stream));
        TOB_NCCL(ncclRecv(v + (long long)(g.nzl() - 1) * g.pstride, cnt, ncclDouble, r + 1, c->comm, c->stream));
    }
    TOB_NCCL(ncclGroupEnd());
    return 0;
}

// Element field with local layers: owned [own0,own1); recv_lo layers below own0 and recv_hi above
// own1 are filled from the z-neighbours; send_down / send_up are the numbers of my boundary
// layers the lower / upper neighbour stores as its ghosts (known analytically by the caller).
int halo_elems(tob_ctx_impl* c, double* v, long long layer_stride, int own0, int own1, int recv_lo, int recv_hi,
               int send_down, int send_up) {
    if (c->desc.nranks == 1) return 0;
    const int r = c->desc.rank, n = c->desc.nranks;
    if (send_down > own1 - own0 || send_up > own1 - own0) TOB_FAIL("ghost width exceeds the slab thickness");
    TOB_NCCL(ncclGroupStart());
    if (r > 0) {
        if (send_down) TOB_NCCL(ncclSend(v + (long long)own0 * layer_stride, (size_t)send_down * layer_stride, ncclDouble, r - 1, c->comm, c->stream));
        if (recv_lo) TOB_NCCL(ncclRecv(v + (long long)(own0 - recv_lo) * layer_stride, (size_t)recv_lo * layer_stride, ncclDouble, r - 1, c->comm, c->stream));
    }
    if (r < n - 1) {
        if (send_up) TOB_NCCL(ncclSend(v + (long long)(own1 - send_up) * layer_stride, (size_t)send_up * layer_stride, ncclDouble, r + 1, c->comm, c->stream));
        if (recv_hi) TOB_NCCL(ncclRecv(v + (long long)own1 * layer_stride, (size_t)recv_hi * layer_stride, ncclDouble, r + 1, c->comm, c->stream));
    }
    TOB_NCCL(ncclGroupEnd());
    return 0;
}

// all-gather of one contiguous chunk per rank (different sizes): rank k's chunk lands at dst + off[k] on every rank
// (set-up traffic and the once-per-V-cycle gather of the coarse residual; NCCL is graph-capturable)
int allgatherv_bytes(tob_ctx_impl* c, const void* mine, void* dst, const size_t* off, const size_t* cnt) {
    const int n = c->desc.nranks, r = c->desc.rank;
    if (n == 1) {
        if (cnt[0]) TOB_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off[0], mine, cnt[0], cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    TOB_NCCL(ncclGroupStart());
    for (int k = 0; k < n; k++) {
        char* recv = static_cast<char*>(dst) + off[k];
        TOB_NCCL(ncclBroadcast(k == r ? mine : static_cast<const void*>(recv), recv, cnt[k], ncclChar, k, c->comm, c->stream));
    }
    TOB_NCCL(ncclGroupEnd());
    return 0;
}

// in-place all-reduce of device scalars
static int allreduce(tob_ctx_impl* c, double* d_vals, int n, int op) {
    if (c->desc.nranks == 1) return 0;
    PeerComm& pc = c->peer;
    if (pc.enabled && n <= kCgSlot) {  // doubles kCgSlot.. of a slot belong to the single-reduction PCG (cgsr.cuh)
        PeerPtrs pp;
        for (int k = 0; k < kMaxRanks; k++) pp.p[k] = k < c->desc.nranks ? pc.peer[k] : nullptr;
        k_finalize_p2p<<<1, 256, 0, c->stream>>>(nullptr, 0, d_vals, n, op, c->desc.nranks, c->desc.rank, pc.arena, pp);
        c->launches++;
        TOB_KERNEL_CHECK();
        return 0;
    }
    TOB_NCCL(ncclAllReduce(d_vals, d_vals, n, ncclDouble, op ? ncclMax : ncclSum, c->comm, c->stream));
    return 0;
}
int allreduce_sum(tob_ctx_impl* c, double* d_vals, int n) { return allreduce(c, d_vals, n, 0); }
int allreduce_max(tob_ctx_impl* c, double* d_vals, int n) { return allreduce(c, d_vals, n, 1); }

// per-CTA partials -> d_out, all-reduced: one launch on the peer path
int finalize_allreduce(tob_ctx_impl* c, int nblocks, int nvals, double* d_out, int op) {
    PeerComm& pc = c->peer;
    if (c->desc.nranks > 1 && pc.enabled && nvals <= kCgSlot) {
        PeerPtrs pp;
        for (int k = 0; k < kMaxRanks; k++) pp.p[k] = k < c->desc.nranks ? pc.peer[k] : nullptr;
        k_finalize_p2p<<<1, 256, 0, c->stream>>>(c->d_partials, nblocks, d_out, nvals, op, c->desc.nranks, c->desc.rank, pc.arena, pp);
        c->launches++;
        TOB_KERNEL_CHECK();
        return 0;
    }
    return -1;  // caller falls back to finalize kernel + allreduce
}

}  // namespace tob
