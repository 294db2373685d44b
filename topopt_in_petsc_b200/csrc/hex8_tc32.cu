// hex8_tc32.cu -- single-precision matrix-free hex8 operator on the 5th-generation tensor cores (tcgen05, TF32 inputs,
// fp32 accumulation in tensor memory): the variant of the operator meant for multigrid SMOOTHING inside an fp64 outer
// Krylov iteration (north_star: "tensor cores used only for the dense 24x24 element contraction"; BASELINE configs[4]:
// "mixed-precision smoother").  The fp64 operator cannot use tensor cores on B200 (tcgen05 has no f64 kind, DMMA shares
// the DFMA pipe, profiles/r01_spmv_probe_v3.json); in fp32 the element contraction IS a tensor-core GEMM:
//
//     Y[128 elements x 24] = V[128 x 24] * KE^T[24 x 24],     V_e = E_e * (nodal values of element e, fixed dofs zeroed)
//
// one tcgen05.mma.kind::tf32 tile of M = 128 (elements), N = 32 (24 + padding), K = 24 = 3 x 8.
//   * CTA = 128 threads = one 16 x 8 patch of elements, marching along z; thread t owns element t of the patch;
//   * the nodal planes (17 x 9 x 3 floats) live in shared memory, the thread gathers its 24 values, scales them by the
//     modulus and stores its ROW of the A operand (K-major canonical layout, no swizzle: 6 x 16-byte chunks);
//   * one elected thread issues the three MMAs (A and B from shared memory, D in TMEM) and commits to an mbarrier;
//   * every thread reads its element's row of D back with tcgen05.ld (lane = element) and adds the 24 values into the
//     two nodal accumulator planes in shared memory (red.shared.add.f32); a plane is complete when the element layers
//     below and above it are done and is written out once (the 15 x 7 interior nodes of the patch: patches overlap by one
//     element, like the fp64 kernel's tiles -- no global atomics).
// SASS: UTCMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit) -- scripts/sass_evidence.sh.
#include "hex8.cuh"
#include "tma.cuh"
#include "vecops.cuh"

namespace tob {

constexpr int TC_EX = 16, TC_EY = 8;            // elements per patch layer
constexpr int TC_NX = TC_EX + 1, TC_NY = TC_EY + 1;
constexpr int TC_PLANE = TC_NX * TC_NY * 3;     // floats per nodal plane of the patch
constexpr int TC_A_BYTES = 6 * 128 * 16;        // A operand: 6 K-chunks x 128 rows x 16 B
constexpr int TC_B_BYTES = 6 * 32 * 16;         // B operand: 6 K-chunks x 32 rows x 16 B

struct Tc32Params {
    const float* u;            // stored nodal vector, fp32, same [z][c][y][x] layout as the fp64 vectors
    float* y;
    const double* E;           // stored element modulus (fp64 as everywhere; converted on load)
    const unsigned char* mask;
    const float* keT;          // [32][24] fp32: row n = KE row n (rows 24..31 zero)
    int nxn, nyn, nzl, px, pe, pm;
    long long nstride, pstride, estride, mstride;
    int zout0, zout1, zchunk;
};

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor (mma_sm100_desc.hpp): start address [0,14) >> 4, LBO [16,30) >> 4, SBO [32,46) >> 4,
    // version [46,48) = 1 on sm_100, base offset 0, layout type [61,64) = 0 (no swizzle, "interleave")
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(128)
k_hex8_tc32(const __grid_constant__ Tc32Params P) {
    __shared__ __align__(128) unsigned char sA[TC_A_BYTES];
    __shared__ __align__(128) unsigned char sB[TC_B_BYTES];
    __shared__ float sU[2][TC_PLANE];    // nodal values, planes e (lower) and e + 1 (upper) of the current element layer
    __shared__ float sAcc[2][TC_PLANE];  // accumulators of the same two planes
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, warp = t >> 5;
    const int i0 = blockIdx.x * (TC_EX - 1) - 1, j0 = blockIdx.y * (TC_EY - 1) - 1;  // node index of the patch corner
    const int z0 = P.zout0 + blockIdx.z * P.zchunk, z1 = min(z0 + P.zchunk, P.zout1);
    if (z0 >= z1) return;

    // ---- one-time set-up: B operand, mbarrier, tensor memory ---------------------------------------------------------------
    for (int idx = t; idx < 32 * 24; idx += 128) {
        const int n = idx / 24, k = idx % 24;
        reinterpret_cast<float*>(sB)[(k >> 2) * (32 * 4) + n * 4 + (k & 3)] = P.keT[idx];
    }
    if (t == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    const uint64_t descA0 = tc_smem_desc(smem_u32(sA), 128 * 16, 8 * 16);
    const uint64_t descB0 = tc_smem_desc(smem_u32(sB), 32 * 16, 8 * 16);
    // cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N = 32 (4 << 17), M = 128 (8 << 24)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (4u << 17) | (8u << 24);

    const int ex = t & (TC_EX - 1), ey = t >> 4;  // my element inside the patch
    const int ei = i0 + ex, ej = j0 + ey;         // its global element index
    const bool elem_in = ei >= 0 && ei < P.nxn - 1 && ej >= 0 && ej < P.nyn - 1;

    // load a nodal plane of the patch (fixed dofs and everything outside the grid as zero), clear its accumulator
    auto load_plane = [&](int buf, int pz) {
        for (int idx = t; idx < TC_PLANE; idx += 128) {
            const int c = idx / (TC_NX * TC_NY), r = idx % (TC_NX * TC_NY), jy = r / TC_NX, ix = r % TC_NX;
            const int gi = i0 + ix, gj = j0 + jy;
            float v = 0.0f;
            if (pz >= 0 && pz < P.nzl && gi >= 0 && gi < P.nxn && gj >= 0 && gj < P.nyn) {
                const unsigned char m = P.mask[(long long)pz * P.mstride + (long long)gj * P.pm + gi];
                if (!((m >> c) & 1)) v = P.u[(long long)pz * P.pstride + c * P.nstride + (long long)gj * P.px + gi];
            }
            sU[buf][idx] = v;
            sAcc[buf][idx] = 0.0f;
        }
    };

    int lo = 0;  // buffer holding the lower plane of the current layer
    load_plane(lo, z0 - 1);
    uint32_t phase = 0;
    for (int e = z0 - 1; e < z1; e++) {  // element layer e lies between node planes e and e + 1
        const int hi = lo ^ 1;
        load_plane(hi, e + 1);
        __syncthreads();
        // ---- my row of the A operand: 24 values (reference dof order 3 * node + c, LinearElasticity.cc:118-120) x modulus ----
        {
            float Ee = 0.0f;
            if (elem_in && e >= 0 && e < P.nzl - 1) Ee = (float)P.E[(long long)e * P.estride + (long long)ej * P.pe + ei];
            float v[24];
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const int a = (n == 1 || n == 2 || n == 5 || n == 6), b = (n == 2 || n == 3 || n == 6 || n == 7), g = n >> 2;
                const float* pl = g ? sU[hi] : sU[lo];
                const int r = (ey + b) * TC_NX + ex + a;
#pragma unroll
                for (int c = 0; c < 3; c++) v[3 * n + c] = Ee * pl[c * (TC_NX * TC_NY) + r];
            }
#pragma unroll
            for (int kc = 0; kc < 6; kc++)
                *reinterpret_cast<float4*>(sA + kc * (128 * 16) + t * 16) = make_float4(v[4 * kc], v[4 * kc + 1], v[4 * kc + 2], v[4 * kc + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's async proxy
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < 3; ks++) {  // K = 24 = 3 steps of 8 (two 16-byte chunks each)
                const uint64_t da = descA0 + (uint64_t)((ks * 2 * 128 * 16) >> 4);
                const uint64_t db = descB0 + (uint64_t)((ks * 2 * 32 * 16) >> 4);
                const uint32_t acc = ks > 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        mbar_wait(&mbar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- my element's row of D: 32 consecutive columns of TMEM lane t ---------------------------------------------------
        uint32_t d[32];
        {
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
                  "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),
                  "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),
                  "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        // ---- scatter-add into the two accumulator planes ---------------------------------------------------------------------
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const int a = (n == 1 || n == 2 || n == 5 || n == 6), b = (n == 2 || n == 3 || n == 6 || n == 7), g = n >> 2;
            float* pl = g ? sAcc[hi] : sAcc[lo];
            const int r = (ey + b) * TC_NX + ex + a;
#pragma unroll
            for (int c = 0; c < 3; c++) atomicAdd(&pl[c * (TC_NX * TC_NY) + r], __uint_as_float(d[3 * n + c]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- plane e is complete: write the interior nodes of the patch -----------------------------------------------------
        if (e >= z0) {
            for (int idx = t; idx < (TC_NX - 2) * (TC_NY - 2) * 3; idx += 128) {
                const int c = idx / ((TC_NX - 2) * (TC_NY - 2)), r = idx % ((TC_NX - 2) * (TC_NY - 2));
                const int jy = 1 + r / (TC_NX - 2), ix = 1 + r % (TC_NX - 2);
                const int gi = i0 + ix, gj = j0 + jy;
                if (gi < P.nxn && gj < P.nyn) {
                    const unsigned char m = P.mask[(long long)e * P.mstride + (long long)gj * P.pm + gi];
                    const long long o = (long long)e * P.pstride + c * P.nstride + (long long)gj * P.px + gi;
                    P.y[o] = ((m >> c) & 1) ? P.u[o] : sAcc[lo][c * (TC_NX * TC_NY) + jy * TC_NX + ix];
                }
            }
        }
        __syncthreads();  // the lower buffers are recycled as the next upper plane
        lo = hi;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

__global__ void __launch_bounds__(kFlatThreads) k_d2f(const double* __restrict__ a, float* __restrict__ b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) b[i] = (float)a[i];
}
__global__ void __launch_bounds__(kFlatThreads) k_f2d(const float* __restrict__ a, double* __restrict__ b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) b[i] = (double)a[i];
}
int vec_d2f(tob_ctx_impl* c, const double* a, float* b, long long n) {
    k_d2f<<<flat_grid(c, n), kFlatThreads, 0, c->stream>>>(a, b, n);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}
int vec_f2d(tob_ctx_impl* c, const float* a, double* b, long long n) {
    k_f2d<<<flat_grid(c, n), kFlatThreads, 0, c->stream>>>(a, b, n);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// y32 = N.(K (N.u32)) + (1-N).u32 on the owned planes; u32 / y32 in the stored layout (ghost planes of u32 current)
int hex8_spmv_tc32(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const float* u32, float* y32, const double* E,
                   const unsigned char* mask, float** keT_cache) {
    if (!*keT_cache) {
        std::vector<float> k(32 * 24, 0.0f);
        for (int n = 0; n < 24; n++)
            for (int kk = 0; kk < 24; kk++) k[n * 24 + kk] = (float)op.KE[n * 24 + kk];
        TOB_CUDA(cudaMalloc(keT_cache, sizeof(float) * k.size()));
        TOB_CUDA(cudaMemcpy(*keT_cache, k.data(), sizeof(float) * k.size(), cudaMemcpyHostToDevice));
    }
    Tc32Params P;
    P.u = u32; P.y = y32; P.E = E; P.mask = mask; P.keT = *keT_cache;
    P.nxn = g.nxn; P.nyn = g.nyn; P.nzl = g.nzl(); P.px = g.px; P.pe = g.pe; P.pm = g.pm;
    P.nstride = g.nstride; P.pstride = g.pstride; P.estride = g.estride; P.mstride = g.mstride;
    P.zout0 = g.zl_own0(); P.zout1 = g.zl_own1();
    const int gx = div_up(g.nxn, TC_EX - 1), gy = div_up(g.nyn, TC_EY - 1);
    const int nz = P.zout1 - P.zout0;
    // enough CTAs for a few waves at ~8 resident CTAs per SM; every chunk recomputes one layer
    int nch = std::max(1, std::min(nz / 8, (int)((8LL * c->num_sms * 3) / std::max(1, gx * gy))));
    P.zchunk = div_up(nz, nch);
    nch = div_up(nz, P.zchunk);
    k_hex8_tc32<<<dim3(gx, gy, nch), 128, 0, c->stream>>>(P);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

}  // namespace tob
