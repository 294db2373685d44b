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


// =========================================================================================================================
// Generation 2: no gather, no scatter -- the tensor core reads the NODAL planes directly.
//
// Store the fp32 vector node-interleaved ("V4": one float4 = (ux, uy, uz, 0) per node, [z][y][x]).  A TMA box of
// (8+2) x (16+2) nodes of one plane lands in shared memory as [18][10] float4.  For a fixed corner offset (a', b', g') the
// values "node I + offset" of the 8 x-consecutive nodes I form exactly one canonical K-major core matrix (8 rows x 16 B,
// rows 16 B apart), row groups (next y) are one tile row = 160 B apart: so the A operand of
//     D_s[I][c] = sum_{a',b',g',c'} u[I - s + (a',b',g')][c'] * KE[(s,c), ((a',b',g'),c')]            (s = (a,b,g): element I - s)
// is a plain UMMA shared-memory descriptor into the plane tile: start = tile + ((1-b+b')*10 + (1-a)) * 16 B, LBO = 16 B
// (the a' = 0 / 1 chunks of a K = 8 step OVERLAP in memory), SBO = 160 B.  The modulus cannot be folded into the operand
// (it belongs to the element, not to the node), so the eight per-element products stay separate TMEM accumulators
// (8 x 16 columns) and each thread -- one node -- finishes  y_I = sum_s E_{I-s} D_s[I]  with 24 FMAs, adds the Chebyshev
// update if asked, and writes its node once.  Per output plane of 128 nodes: 32 tcgen05.mma (M 128, N 16, K 8), one commit,
// 8 tcgen05.ld per thread; shared memory sees only the TMA writes.  No CUDA-core gather/scatter, no atomics, no tile overlap.
// =========================================================================================================================
constexpr int T2_X = 8, T2_Y = 16;                 // output nodes per plane per CTA (= 128 = UMMA M)
constexpr int T2_PX = T2_X + 2, T2_PY = T2_Y + 2;  // plane tile with halo
constexpr int T2_SLOT = ((T2_PX * T2_PY * 16 + 127) / 128) * 128;
constexpr int T2_B_BYTES = 32 * 512;               // 8 shifts x 4 K-steps x (2 chunks x 16 rows x 16 B)

struct Tc2Params {
    float4* y;                 // V4 output (EPI 0: A x; EPI 1: x + d)
    const float* E32;          // element modulus, fp32, stored element layout
    const unsigned char* mask;
    const float* B;            // [32][2][16][4] operand blocks (host-prepared from KE)
    const float4* b;           // EPI 1: right-hand side
    const float4* dinv;        // EPI 1: 1/diag
    float4* d;                 // EPI 1: Chebyshev direction (in/out)
    float s0, s1;
    int nxn, nyn, nzl, pe, pm;
    long long estride, mstride;
    long long p4;              // nodes per V4 plane (nxn * nyn rounded up to even: a plane is a whole number of 32-byte units for the halo kernels)
    int zout0, zout1, zchunk;
    int epi;
};

// PACKED: the eight accumulators share one 64-column TMEM window (D_s at column 4 s; the first MMA of a plane is an N = 32
// tile that also clears the window, every other MMA accumulates -- rows 3..15 of the operand blocks are zero, so they add
// nothing to their neighbours' columns) and the epilogue needs ONE tcgen05.ld instead of eight; 64 instead of 128 columns
// also doubles the CTAs an SM can hold.
template <bool PACKED>
__global__ void __launch_bounds__(128)
k_hex8_tc32b(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ Tc2Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;                    // operand blocks of KE
    unsigned char* sP = smem + T2_B_BYTES + 1024;  // 4 plane slots (after the N = 32 clearing block)
    __shared__ __align__(8) uint64_t full[4];
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, warp = t >> 5;
    const int ix = t & (T2_X - 1), iy = t >> 3;
    const int gi = blockIdx.x * T2_X + ix, gj = blockIdx.y * T2_Y + iy;
    const int z0 = P.zout0 + blockIdx.z * P.zchunk, z1 = min(z0 + P.zchunk, P.zout1);
    if (z0 >= z1) return;

    for (int idx = t; idx < (T2_B_BYTES + 1024) / 16; idx += 128) reinterpret_cast<float4*>(sB)[idx] = reinterpret_cast<const float4*>(P.B)[idx];
    if (t == 0) {
        for (int s = 0; s < 4; s++) mbar_init(&full[s], 1);
        mbar_init(&mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(PACKED ? 64 : 128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // sB was written through the generic proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // D = F32, A = B = TF32, K-major both, N = 16, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (2u << 17) | (8u << 24);
    const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | (4u << 17) | (8u << 24);  // N = 32
    const uint32_t sB_addr = smem_u32(sB), sP_addr = smem_u32(sP);

    auto issue_plane = [&](int pz) {  // TMA: plane pz of this tile (with halo) -> slot pz & 3 (outside the grid: zeros)
        const int s = pz & 3;
        mbar_expect_tx(&full[s], T2_PX * T2_PY * 16);
        tma_load_4d(sP + s * T2_SLOT, &tmX, &full[s], 0, (int)blockIdx.x * T2_X - 1, (int)blockIdx.y * T2_Y - 1, pz);
    };
    if (t == 0) {
        for (int pz = z0 - 1; pz <= z0 + 2 && pz <= z1; pz++) issue_plane(pz);
    }
    // modulus of the 4 elements (ix - a, iy - b) of a layer
    auto load_E = [&](int layer, float (&Ev)[4]) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int ei = gi - (q & 1), ej = gj - (q >> 1);
            Ev[q] = (layer >= 0 && layer < P.nzl - 1 && ei >= 0 && ei < P.nxn - 1 && ej >= 0 && ej < P.nyn - 1)
                        ? P.E32[(long long)layer * P.estride + (long long)ej * P.pe + ei] : 0.0f;
        }
    };
    float Elo[4], Ehi[4];  // layers z - 1 (g = 1) and z (g = 0)
    load_E(z0 - 1, Elo);
    uint32_t mphase = 0;
    for (int z = z0; z < z1; z++) {
        if (t == 0) {
            // planes z - 1, z, z + 1 must have landed (phase of a slot: one completion per 4 planes)
            for (int pz = z - 1; pz <= z + 1; pz++) mbar_wait(&full[pz & 3], (uint32_t)(((pz - (z0 - 1)) >> 2) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const int a = s & 1, b = (s >> 1) & 1, g = s >> 2;
#pragma unroll
                for (int st = 0; st < 4; st++) {
                    const int bp = st & 1, gp = st >> 1;
                    const int pz = z - g + gp;
                    const uint32_t aaddr = sP_addr + (uint32_t)((pz & 3) * T2_SLOT + ((1 - b + bp) * T2_PX + (1 - a)) * 16);
                    const uint64_t da = tc_smem_desc(aaddr, 16, T2_PX * 16);
                    const bool first = PACKED && s == 0 && st == 0;  // N = 32 block (rows 3..31 zero): writes D_0 and clears the window
                    const uint64_t db = first ? tc_smem_desc(sB_addr + (uint32_t)T2_B_BYTES, 512, 128)
                                              : tc_smem_desc(sB_addr + (uint32_t)((s * 4 + st) * 512), 256, 128);
                    const uint32_t acc = PACKED ? (first ? 0u : 1u) : (uint32_t)(st > 0);
                    const uint32_t id = first ? idesc32 : idesc;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem + (uint32_t)((PACKED ? 4 : 16) * s)), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done)) : "memory");
        }
        load_E(z, Ehi);
        // my own nodal value (Dirichlet pass-through / Chebyshev update) sits in the plane tile; read it before the slot is recycled
        float4 xo;
        {
            // the tile of plane z has landed for thread 0 only; everybody else learns it through mma_done below, so read afterwards
        }
        mbar_wait(&mma_done, mphase);
        mphase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mbar_wait(&full[z & 3], (uint32_t)(((z - (z0 - 1)) >> 2) & 1));  // (already complete: makes the TMA writes of plane z visible to me)
        xo = *reinterpret_cast<const float4*>(sP + (z & 3) * T2_SLOT + ((iy + 1) * T2_PX + ix + 1) * 16);
        float acc[3] = {0.0f, 0.0f, 0.0f};
        if (PACKED) {
            uint32_t d[32];
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
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const float Es = (s >> 2) ? Elo[s & 3] : Ehi[s & 3];
                acc[0] = fmaf(Es, __uint_as_float(d[4 * s]), acc[0]);
                acc[1] = fmaf(Es, __uint_as_float(d[4 * s + 1]), acc[1]);
                acc[2] = fmaf(Es, __uint_as_float(d[4 * s + 2]), acc[2]);
            }
        } else {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                uint32_t r0, r1, r2, r3;
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(16 * s);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float Es = (s >> 2) ? Elo[s & 3] : Ehi[s & 3];
                acc[0] = fmaf(Es, __uint_as_float(r0), acc[0]);
                acc[1] = fmaf(Es, __uint_as_float(r1), acc[1]);
                acc[2] = fmaf(Es, __uint_as_float(r2), acc[2]);
                (void)r3;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (gi < P.nxn && gj < P.nyn) {
            const unsigned char m = P.mask[(long long)z * P.mstride + (long long)gj * P.pm + gi];
            const float xs[3] = {xo.x, xo.y, xo.z};
            float au[3];
#pragma unroll
            for (int c = 0; c < 3; c++) au[c] = ((m >> c) & 1) ? xs[c] : acc[c];
            const long long o = (long long)z * P.p4 + (long long)gj * P.nxn + gi;
            if (P.epi == 0) {
                P.y[o] = make_float4(au[0], au[1], au[2], 0.0f);
            } else {
                const float4 bv = P.b[o], dv = P.dinv[o], dd = P.d[o];
                const float d0 = P.s0 * dd.x + P.s1 * dv.x * (bv.x - au[0]);
                const float d1 = P.s0 * dd.y + P.s1 * dv.y * (bv.y - au[1]);
                const float d2 = P.s0 * dd.z + P.s1 * dv.z * (bv.z - au[2]);
                P.d[o] = make_float4(d0, d1, d2, 0.0f);
                P.y[o] = make_float4(xo.x + d0, xo.y + d1, xo.z + d2, 0.0f);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) Elo[q] = Ehi[q];
        __syncthreads();  // every thread is done with TMEM and with the tile of plane z - 1: recycle its slot
        if (t == 0 && z + 3 <= z1) issue_plane(z + 3);
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(PACKED ? 64 : 128) : "memory");
}

// ---- V4 (node-interleaved fp32) <-> stored fp64 layout ----------------------------------------------------------------------
long long v4_plane_nodes(const LevelGrid& g) { return ((long long)g.nxn * g.nyn + 1) & ~1LL; }
__global__ void k_st_to_v4(const double* __restrict__ st, float4* __restrict__ v4, int nxn, int nyn, int nzl, int px, long long nstride,
                           long long pstride, long long p4) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)nxn * nyn * nzl) return;
    const int i = n % nxn, j = (n / nxn) % nyn, k = n / ((long long)nxn * nyn);
    const long long o = (long long)k * pstride + (long long)j * px + i;
    v4[(long long)k * p4 + (long long)j * nxn + i] = make_float4((float)st[o], (float)st[o + nstride], (float)st[o + 2 * nstride], 0.0f);
}
__global__ void k_v4_to_st(const float4* __restrict__ v4, double* __restrict__ st, int nxn, int nyn, int nzl, int px, long long nstride,
                           long long pstride, int z0, int z1, long long p4) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)nxn * nyn * nzl) return;
    const int i = n % nxn, j = (n / nxn) % nyn, k = n / ((long long)nxn * nyn);
    if (k < z0 || k >= z1) return;
    const long long o = (long long)k * pstride + (long long)j * px + i;
    const float4 v = v4[(long long)k * p4 + (long long)j * nxn + i];
    st[o] = v.x; st[o + nstride] = v.y; st[o + 2 * nstride] = v.z;
}
int vec_st_to_v4(tob_ctx_impl* c, const LevelGrid& g, const double* st, float4* v4) {
    const long long n = (long long)g.nxn * g.nyn * g.nzl();
    k_st_to_v4<<<div_up(n, 256), 256, 0, c->stream>>>(st, v4, g.nxn, g.nyn, g.nzl(), g.px, g.nstride, g.pstride, v4_plane_nodes(g));
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}
int vec_v4_to_st(tob_ctx_impl* c, const LevelGrid& g, const float4* v4, double* st) {
    const long long n = (long long)g.nxn * g.nyn * g.nzl();
    k_v4_to_st<<<div_up(n, 256), 256, 0, c->stream>>>(v4, st, g.nxn, g.nyn, g.nzl(), g.px, g.nstride, g.pstride, g.zl_own0(), g.zl_own1(), v4_plane_nodes(g));
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// operand blocks of KE for the 8 shifts x 4 K-steps: B[(s*4+st)][chunk a'][row n (16)][4] = KE[3 ref(s) + n][3 ref(a',b',g') + c']
int tc2_build_B(const Hex8Operator& op, float** dB) {
    if (*dB) return 0;
    auto ref = [](int a, int b, int g) { return (g ? 4 : 0) + (b ? (a ? 2 : 3) : (a ? 1 : 0)); };
    std::vector<float> B(32 * 2 * 16 * 4 + 256, 0.0f);  // + the N = 32 clearing block of (s, st) = (0, 0): [chunk 2][32 rows][4]
    for (int s = 0; s < 8; s++)
        for (int st = 0; st < 4; st++)
            for (int ap = 0; ap < 2; ap++)
                for (int n = 0; n < 3; n++)
                    for (int cp = 0; cp < 3; cp++) {
                        const int row = 3 * ref(s & 1, (s >> 1) & 1, s >> 2) + n;
                        const int col = 3 * ref(ap, st & 1, st >> 1) + cp;
                        B[(size_t)(((s * 4 + st) * 2 + ap) * 16 + n) * 4 + cp] = (float)op.KE[row * 24 + col];
                    }
    for (int ap = 0; ap < 2; ap++)
        for (int n = 0; n < 3; n++)
            for (int cp = 0; cp < 3; cp++)
                B[(size_t)32 * 2 * 16 * 4 + (size_t)(ap * 32 + n) * 4 + cp] = (float)op.KE[(3 * ref(0, 0, 0) + n) * 24 + 3 * ref(ap, 0, 0) + cp];
    TOB_CUDA(cudaMalloc(dB, sizeof(float) * B.size()));
    TOB_CUDA(cudaMemcpy(*dB, B.data(), sizeof(float) * B.size(), cudaMemcpyHostToDevice));
    return 0;
}

// y4 = A x4 (epi 0) or one Chebyshev step (epi 1) on V4 vectors (ghost planes of x4 current, fixed dofs of x4 zero on entry of epi 0
// is NOT required: fixed dofs are passed through, but they must not feed the products, so the caller keeps them zero)
int hex8_tc32b(tob_ctx_impl* c, const LevelGrid& g, const float* dB, const float4* x4, float4* y4, const float* E32, const unsigned char* mask,
               int epi, const float4* b4, const float4* dinv4, float4* d4, float s0, float s1) {
    CUtensorMap tmX;
    {
        cuuint64_t dims[4] = {4, (cuuint64_t)g.nxn, (cuuint64_t)g.nyn, (cuuint64_t)g.nzl()};
        cuuint64_t str[3] = {16, (cuuint64_t)g.nxn * 16, (cuuint64_t)v4_plane_nodes(g) * 16};
        cuuint32_t box[4] = {4, T2_PX, T2_PY, 1};
        TOB_CHECK(encode_map(c, &tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x4, dims, str, box));
    }
    Tc2Params P;
    P.y = y4; P.E32 = E32; P.mask = mask; P.B = dB; P.b = b4; P.dinv = dinv4; P.d = d4; P.s0 = s0; P.s1 = s1;
    P.nxn = g.nxn; P.nyn = g.nyn; P.nzl = g.nzl(); P.pe = g.pe; P.pm = g.pm; P.estride = g.estride; P.mstride = g.mstride;
    P.zout0 = g.zl_own0(); P.zout1 = g.zl_own1(); P.epi = epi; P.p4 = v4_plane_nodes(g);
    const int gx = div_up(g.nxn, T2_X), gy = div_up(g.nyn, T2_Y);
    const int nz = P.zout1 - P.zout0;
    int nch = std::max(1, std::min(nz / 8, (int)((6LL * c->num_sms * 2) / std::max(1, gx * gy))));
    P.zchunk = div_up(nz, nch);
    nch = div_up(nz, P.zchunk);
    const int smem_bytes = T2_B_BYTES + 1024 + 4 * T2_SLOT;
    static const bool packed = !(getenv("TOB_TC32_PACKED") && atoi(getenv("TOB_TC32_PACKED")) == 0);
    if (packed) k_hex8_tc32b<true><<<dim3(gx, gy, nch), 128, smem_bytes, c->stream>>>(tmX, P);
    else k_hex8_tc32b<false><<<dim3(gx, gy, nch), 128, smem_bytes, c->stream>>>(tmX, P);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}


// =========================================================================================================================
// Generation 3: wide tiles.  ncu on generation 2 (profiles/r02_ncu_tc32b_256.txt): the tensor pipe is busy 79 % of the time --
// an M128 x K8 tcgen05.mma costs ~33 cycles however narrow N is, and the per-shift products are N = 3.  Here the rows are
// ELEMENTS (the window with lower corner node I), so one tile computes all 24 outputs of the element at once:
//     D[e][0..23] = KE * u_e            4 MMAs of M 128, N 32, K 8 per 128 elements  (instead of 32 of N 16)
// still without any gather (the A operand is the overlapping-descriptor view of the raw nodal planes).  What comes back is
// per ELEMENT, so the epilogue scales the row by the element's modulus and sums the eight element contributions of a node:
// the four elements of a layer around a node are lane - 1 (x), lane - 8 (y) and lane - 9 of the same warp, so the sums are
// warp shuffles; only the element row that crosses a warp boundary goes through shared memory (YM 2, the default; YM 0 / 1 keep
// the earlier transposed-buffer forms for A/B).  Contributions to the plane above are carried in registers to the next layer.
// Tiles overlap by one element (7 x 15 complete nodes per 8 x 16 elements).  The MMAs of layer L + 1 are issued before the node
// sums of layer L (two TMEM accumulators), so the tensor pipe works under the epilogue.  8 CTAs per SM (the TMEM limit).
// =========================================================================================================================
constexpr int T3_EX = 8, T3_EY = 16;               // elements per tile layer (= UMMA M)
constexpr int T3_PX = T3_EX + 1, T3_PY = T3_EY + 1;
constexpr int T3_SLOT = ((T3_PX * T3_PY * 16 + 127) / 128) * 128;
constexpr int T3_B_BYTES = 4 * 1024;               // 4 K-steps x (2 chunks x 32 rows x 16 B)
constexpr int T3_Y_BYTES = 8 * 128 * 16;           // one transposed output buffer: [local node (8)][element (128)] float4
constexpr int T3_X_BYTES = 2 * 4 * 6 * 8 * 4;      // YM 2: [buffer][warp][6 values][8 lanes] floats crossing a warp boundary

// Y3: the transposed buffer holds three float planes per local node ([24][128] floats, 12 KB, conflict-free 32-bit accesses: 8 CTAs
// per SM) instead of one float4 ([8][128], 16 KB, a quarter of the wavefronts carry the padding lane: 7 CTAs per SM)
// YM 2: no transposed buffer at all -- the node sums are formed with warp shuffles (x neighbour = lane - 1, y neighbour = lane - 8);
// only the row that crosses a warp boundary (6 floats of 8 lanes per warp) goes through shared memory, double-buffered, so a layer
// needs ONE block barrier and 18 shuffles + 12 quarter-warp shared accesses per warp instead of 48 full-warp shared accesses
template <int YM>
__global__ void __launch_bounds__(128)
k_hex8_tc32c(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ Tc2Params P) {
    extern __shared__ __align__(128) unsigned char smem[];  // (no swizzle: the operand descriptors need 16-byte alignment only)
    unsigned char* sB = smem;
    unsigned char* sP = smem + T3_B_BYTES;                     // 4 plane slots
    float4* sY = reinterpret_cast<float4*>(smem + T3_B_BYTES + 4 * T3_SLOT);  // transposed outputs (one buffer: occupancy beats the saved barrier)
    // barriers and the TMEM slot live behind the buffers (static shared memory would round the CTA up by 2 KB: 7 instead of 8 per SM)
    constexpr bool Y3 = (YM == 1);
    unsigned char* tail = smem + T3_B_BYTES + 4 * T3_SLOT + (YM == 2 ? T3_X_BYTES : YM == 1 ? T3_Y_BYTES * 3 / 4 : T3_Y_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);          // [4]
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(tail + 32); // [2]: one per TMEM accumulator (the MMAs run one layer ahead of the epilogue)
    uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(tail + 48);

    const int t = threadIdx.x, warp = t >> 5;
    const int ix = t & (T3_EX - 1), iy = t >> 3;
    const int oi = blockIdx.x * (T3_EX - 1) - 1, oj = blockIdx.y * (T3_EY - 1) - 1;  // node index of the tile origin
    const int gi = oi + ix, gj = oj + iy;  // my element (lower-corner node) = my node
    const int z0 = P.zout0 + blockIdx.z * P.zchunk, z1 = min(z0 + P.zchunk, P.zout1);
    if (z0 >= z1) return;

    for (int idx = t; idx < T3_B_BYTES / 16; idx += 128) reinterpret_cast<float4*>(sB)[idx] = reinterpret_cast<const float4*>(P.B)[idx];
    if (t == 0) {
        for (int s = 0; s < 4; s++) mbar_init(&full[s], 1);
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (4u << 17) | (8u << 24);  // F32 += TF32 x TF32, N 32, M 128
    const uint32_t sB_addr = smem_u32(sB), sP_addr = smem_u32(sP);
    const int pfirst = z0 - 1;  // first plane / layer of the chunk

    auto issue_plane = [&](int pz) {
        const int s = pz & 3;
        mbar_expect_tx(&full[s], T3_PX * T3_PY * 16);
        tma_load_4d(sP + s * T3_SLOT, &tmX, &full[s], 0, oi, oj, pz);
    };
    auto wait_plane = [&](int pz) { mbar_wait(&full[pz & 3], (uint32_t)(((pz - pfirst) >> 2) & 1)); };
    auto issue_mma = [&](int L) {  // layer L: planes L (g' = 0) and L + 1 (g' = 1); accumulator (L - pfirst) & 1
        const int buf = (L - pfirst) & 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int st = 0; st < 4; st++) {
            const int bp = st & 1, gp = st >> 1;
            const uint32_t aaddr = sP_addr + (uint32_t)(((L + gp) & 3) * T3_SLOT + bp * T3_PX * 16);
            const uint64_t da = tc_smem_desc(aaddr, 16, T3_PX * 16);
            const uint64_t db = tc_smem_desc(sB_addr + (uint32_t)(st * 1024), 512, 128);
            const uint32_t acc = st > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem + (uint32_t)(32 * buf)), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done[buf])) : "memory");
    };
    if (t == 0) {
        for (int pz = pfirst; pz <= pfirst + 3 && pz <= z1; pz++) issue_plane(pz);
        wait_plane(pfirst);
        wait_plane(pfirst + 1);
        issue_mma(pfirst);
    }
    const bool elem_in = gi >= 0 && gi < P.nxn - 1 && gj >= 0 && gj < P.nyn - 1;
    const bool node_out = ix >= 1 && iy >= 1 && gi < P.nxn && gj < P.nyn;
    float carry[3] = {0.0f, 0.0f, 0.0f};
    for (int L = pfirst; L < z1; L++) {
        const float Ee = (elem_in && L >= 0 && L < P.nzl - 1) ? P.E32[(long long)L * P.estride + (long long)gj * P.pe + gi] : 0.0f;
        if (t == 0 && L + 1 < z1) {  // the tensor core works one layer ahead: its accumulator was read out an iteration ago
            wait_plane(L + 2);       // (plane L + 1 was waited for before the previous MMA)
            issue_mma(L + 1);
        }
        const int buf = (L - pfirst) & 1;
        mbar_wait(&mma_done[buf], (uint32_t)(((L - pfirst) >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float4* Yb = sY;
        float low[3] = {0.0f, 0.0f, 0.0f}, up[3] = {0.0f, 0.0f, 0.0f};
        {
            uint32_t d[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(32 * buf);
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
            if (YM == 2) {
                // D columns are [local node][4]; local node (a, b, g) = g * 4 + (b ? (a ? 2 : 3) : (a ? 1 : 0))
                const int lane = t & 31;
                float* X = reinterpret_cast<float*>(sY) + ((L - pfirst) & 1) * (4 * 6 * 8);
                float a1[6];
#pragma unroll
                for (int g = 0; g < 2; g++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        // x direction: my element's (0, b) node + the (1, b) node of the element to the left (lane - 1)
                        const float v00 = Ee * __uint_as_float(d[4 * (4 * g + 0) + c]), v10 = Ee * __uint_as_float(d[4 * (4 * g + 1) + c]);
                        const float v01 = Ee * __uint_as_float(d[4 * (4 * g + 3) + c]), v11 = Ee * __uint_as_float(d[4 * (4 * g + 2) + c]);
                        const float a0 = v00 + __shfl_up_sync(0xffffffffu, v10, 1);
                        a1[3 * g + c] = v01 + __shfl_up_sync(0xffffffffu, v11, 1);
                        if (g == 0) low[c] = a0; else up[c] = a0;
                    }
                if (lane >= 24 && warp < 3) {
#pragma unroll
                    for (int q = 0; q < 6; q++) X[(warp * 6 + q) * 8 + (lane - 24)] = a1[q];
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();  // crossing rows stored; TMEM free; everybody is past the epilogue of layer L - 1 (slot of plane L - 1 free)
                if (t == 0 && L + 3 <= z1 && L + 3 > pfirst + 3) issue_plane(L + 3);
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    // y direction: the (., 1) sums of the element row below (lane - 8, or the previous warp's last row)
                    float bsum = __shfl_up_sync(0xffffffffu, a1[q], 8);
                    if (lane < 8) bsum = warp > 0 ? X[((warp - 1) * 6 + q) * 8 + lane] : 0.0f;
                    if (q < 3) low[q] += bsum; else up[q - 3] += bsum;
                }
            } else {
#pragma unroll
                for (int n = 0; n < 8; n++) {  // D columns are laid out [local node][4] (the operand blocks pad every node to 4 rows)
                    if (Y3) {
                        float* Yf = reinterpret_cast<float*>(sY);
#pragma unroll
                        for (int c = 0; c < 3; c++) Yf[(3 * n + c) * 128 + t] = Ee * __uint_as_float(d[4 * n + c]);
                    } else {
                        Yb[n * 128 + t] = make_float4(Ee * __uint_as_float(d[4 * n]), Ee * __uint_as_float(d[4 * n + 1]), Ee * __uint_as_float(d[4 * n + 2]), 0.0f);
                    }
                }
            }
        }
        if (YM != 2) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();  // Yb complete; TMEM free; everybody is past the epilogue of layer L - 1 (slot of plane L - 1 free)
            if (t == 0 && L + 3 <= z1 && L + 3 > pfirst + 3) issue_plane(L + 3);
            // ---- node sums: plane L gets `carry` (from layer L - 1) + the lower-node outputs of layer L ---------------------------
            if (ix >= 1 && iy >= 1) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int a = q & 1, b = q >> 1;
                    const int e = t - (8 * b + a);
                    const int nlo = b ? (a ? 2 : 3) : (a ? 1 : 0);  // reference local node (a, b, 0); (a, b, 1) = nlo + 4
                    if (Y3) {
                        const float* Yf = reinterpret_cast<const float*>(sY);
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            low[c] += Yf[(3 * nlo + c) * 128 + e];
                            up[c] += Yf[(3 * (nlo + 4) + c) * 128 + e];
                        }
                    } else {
                        const float4 vl = Yb[nlo * 128 + e], vu = Yb[(nlo + 4) * 128 + e];
                        low[0] += vl.x; low[1] += vl.y; low[2] += vl.z;
                        up[0] += vu.x; up[1] += vu.y; up[2] += vu.z;
                    }
                }
            }
        }
        if (L >= z0 && node_out) {
            wait_plane(L);  // (complete long ago: orders my generic read of the TMA-written tile)
            const float4 xo = *reinterpret_cast<const float4*>(sP + (L & 3) * T3_SLOT + (iy * T3_PX + ix) * 16);
            const unsigned char m = P.mask[(long long)L * P.mstride + (long long)gj * P.pm + gi];
            const float xs[3] = {xo.x, xo.y, xo.z};
            float au[3];
#pragma unroll
            for (int c = 0; c < 3; c++) au[c] = ((m >> c) & 1) ? xs[c] : carry[c] + low[c];
            const long long o = (long long)L * P.p4 + (long long)gj * P.nxn + gi;
            if (P.epi == 0) {
                P.y[o] = make_float4(au[0], au[1], au[2], 0.0f);
            } else {
                const float4 bv = P.b[o], dv = P.dinv[o], dd = P.d[o];
                const float d0 = P.s0 * dd.x + P.s1 * dv.x * (bv.x - au[0]);
                const float d1 = P.s0 * dd.y + P.s1 * dv.y * (bv.y - au[1]);
                const float d2 = P.s0 * dd.z + P.s1 * dv.z * (bv.z - au[2]);
                P.d[o] = make_float4(d0, d1, d2, 0.0f);
                P.y[o] = make_float4(xo.x + d0, xo.y + d1, xo.z + d2, 0.0f);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) carry[c] = up[c];
        if (YM != 2) __syncthreads();  // the node sums are done with the transposed buffer
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

// operand blocks for generation 3: B[step (b', g')][chunk a'][row 4 * node + c (32)][4] = KE[3 * node + c][3 ref(a', b', g') + c']
int tc3_build_B(const Hex8Operator& op, float** dB) {
    if (*dB) return 0;
    auto ref = [](int a, int b, int g) { return (g ? 4 : 0) + (b ? (a ? 2 : 3) : (a ? 1 : 0)); };
    std::vector<float> B(4 * 2 * 32 * 4, 0.0f);
    for (int st = 0; st < 4; st++)
        for (int ap = 0; ap < 2; ap++)
            for (int nd = 0; nd < 8; nd++)      // output row 4 * node + c (c < 3; the fourth row of every node stays zero)
                for (int c = 0; c < 3; c++)
                    for (int cp = 0; cp < 3; cp++)
                        B[(size_t)((st * 2 + ap) * 32 + 4 * nd + c) * 4 + cp] = (float)op.KE[(3 * nd + c) * 24 + 3 * ref(ap, st & 1, st >> 1) + cp];
    TOB_CUDA(cudaMalloc(dB, sizeof(float) * B.size()));
    TOB_CUDA(cudaMemcpy(*dB, B.data(), sizeof(float) * B.size(), cudaMemcpyHostToDevice));
    return 0;
}

int hex8_tc32c(tob_ctx_impl* c, const LevelGrid& g, const float* dB, const float4* x4, float4* y4, const float* E32, const unsigned char* mask,
               int epi, const float4* b4, const float4* dinv4, float4* d4, float s0, float s1) {
    CUtensorMap tmX;
    {
        cuuint64_t dims[4] = {4, (cuuint64_t)g.nxn, (cuuint64_t)g.nyn, (cuuint64_t)g.nzl()};
        cuuint64_t str[3] = {16, (cuuint64_t)g.nxn * 16, (cuuint64_t)v4_plane_nodes(g) * 16};
        cuuint32_t box[4] = {4, T3_PX, T3_PY, 1};
        TOB_CHECK(encode_map(c, &tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x4, dims, str, box));
    }
    Tc2Params P;
    P.y = y4; P.E32 = E32; P.mask = mask; P.B = dB; P.b = b4; P.dinv = dinv4; P.d = d4; P.s0 = s0; P.s1 = s1;
    P.nxn = g.nxn; P.nyn = g.nyn; P.nzl = g.nzl(); P.pe = g.pe; P.pm = g.pm; P.estride = g.estride; P.mstride = g.mstride;
    P.zout0 = g.zl_own0(); P.zout1 = g.zl_own1(); P.epi = epi; P.p4 = v4_plane_nodes(g);
    const int gx = div_up(g.nxn, T3_EX - 1), gy = div_up(g.nyn, T3_EY - 1);
    const int nz = P.zout1 - P.zout0;
    // z-chunks: ~7 CTAs fit an SM (30 KB of shared memory each); pick the chunk count that minimises waves x (planes per chunk + 2)
    static const int ym = getenv("TOB_TC32_YM") ? atoi(getenv("TOB_TC32_YM")) : 2;  // 0 float4 buffer, 1 three float planes, 2 shuffles
    const long long cap = (ym >= 1 ? 8LL : 7LL) * c->num_sms, tiles = (long long)gx * gy;
    int nch = 1;
    {
        double best = 1e300;
        for (int k = 1; k <= 16 && k <= std::max(1, nz / 4); k++) {
            const int zc = div_up(nz, k), kk = div_up(nz, zc);
            const double cost = ceil((double)(tiles * kk) / cap) * (zc + 2.0);
            if (cost < best * 0.999) { best = cost; nch = kk; }
        }
    }
    P.zchunk = div_up(nz, nch);
    nch = div_up(nz, P.zchunk);
    const int smem_bytes = T3_B_BYTES + 4 * T3_SLOT + (ym == 2 ? T3_X_BYTES : ym == 1 ? T3_Y_BYTES * 3 / 4 : T3_Y_BYTES) + 64;
    static bool carve[64] = {};
    if (c->device >= 0 && c->device < 64 && !carve[c->device]) {  // 8 x 27 KB only fit with the largest shared-memory carve-out
        cudaFuncSetAttribute(k_hex8_tc32c<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_hex8_tc32c<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_hex8_tc32c<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carve[c->device] = true;
    }
    if (ym == 2) k_hex8_tc32c<2><<<dim3(gx, gy, nch), 128, smem_bytes, c->stream>>>(tmX, P);
    else if (ym == 1) k_hex8_tc32c<1><<<dim3(gx, gy, nch), 128, smem_bytes, c->stream>>>(tmX, P);
    else k_hex8_tc32c<0><<<dim3(gx, gy, nch), 128, smem_bytes, c->stream>>>(tmX, P);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

}  // namespace tob
