// Stand-alone bisect of TMA tensor-map variants (run one variant per process).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int bytes, double* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(sm)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(sm)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        if (RANK == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(s32(sm)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    if (threadIdx.x == 0) { out[0] = ((double*)sm)[0]; out[1] = ((double*)sm)[1]; out[2] = ((double*)sm)[33]; out[3] = (double)sm[1]; }
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int v = argc > 1 ? atoi(argv[1]) : 0;
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    Enc enc = (Enc)fn;
    const int nx = 36, ny = 21, nz = 13;  // px = 36
    double* d; cudaMalloc(&d, sizeof(double) * nx * ny * 3 * nz);
    double* h = (double*)malloc(sizeof(double) * nx * ny * 3 * nz);
    for (int i = 0; i < nx * ny * 3 * nz; i++) h[i] = i;
    cudaMemcpy(d, h, sizeof(double) * nx * ny * 3 * nz, cudaMemcpyHostToDevice);
    double* out; cudaMalloc(&out, 64);
    CUtensorMap tm; cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = CUDA_SUCCESS; int rank = 2, bytes = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (v == 1 || v == 2 || v == 7) {  // 2D f64 box 32x8 ; v7 coords (1,0) (unaligned inner start)
        cuuint64_t dims[2] = {34, (cuuint64_t)ny}; cuuint64_t str[1] = {nx * 8}; cuuint32_t box[2] = {32, 8};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        rank = 2; bytes = 32 * 8 * 8; if (v == 2) { c0 = -1; c1 = -1; } if (v == 7) { c0 = 1; }
    } else if (v == 3 || v == 4) {  // 4D f64 as in the kernel
        cuuint64_t dims[4] = {34, (cuuint64_t)ny, 3, (cuuint64_t)nz}; cuuint64_t str[3] = {nx * 8, (cuuint64_t)nx * ny * 8, (cuuint64_t)nx * ny * 3 * 8}; cuuint32_t box[4] = {32, 8, 3, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        rank = 4; bytes = 32 * 8 * 3 * 8; if (v == 4) { c0 = -1; c1 = -1; c3 = -1; }
    } else if (v == 5) {  // 2D f64 box inner 16 (128 B)
        cuuint64_t dims[2] = {34, (cuuint64_t)ny}; cuuint64_t str[1] = {nx * 8}; cuuint32_t box[2] = {16, 8};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        rank = 2; bytes = 16 * 8 * 8;
    } else if (v == 6) {  // 3D u8 box 32x8x1
        cuuint64_t dims[3] = {34, (cuuint64_t)ny, (cuuint64_t)nz}; cuuint64_t str[2] = {48, 48 * ny}; cuuint32_t box[3] = {32, 8, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        rank = 3; bytes = 32 * 8;
    } else if (v == 8) {  // 2D treated as UINT64 instead of FLOAT64
        cuuint64_t dims[2] = {34, (cuuint64_t)ny}; cuuint64_t str[1] = {nx * 8}; cuuint32_t box[2] = {32, 8};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        rank = 2; bytes = 32 * 8 * 8; c0 = -1;
    } else if (v == 9) {  // 2D f32 view of the f64 data: inner 64 floats, coords in float units (odd double start => c0 = -2)
        cuuint64_t dims[2] = {68, (cuuint64_t)ny}; cuuint64_t str[1] = {nx * 8}; cuuint32_t box[2] = {64, 8};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        rank = 2; bytes = 64 * 4 * 8; c0 = -2;
    }
    printf("variant %d encode=%d\n", v, (int)r);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    if (rank == 2) k<2><<<1, 32, 16384>>>(tm, c0, c1, c2, c3, bytes, out);
    if (rank == 3) k<3><<<1, 32, 16384>>>(tm, c0, c1, c2, c3, bytes, out);
    if (rank == 4) k<4><<<1, 32, 16384>>>(tm, c0, c1, c2, c3, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    double ho[4] = {0, 0, 0, 0};
    cudaMemcpy(ho, out, 32, cudaMemcpyDeviceToHost);
    printf("variant %d result=%s  sm[0]=%g sm[1]=%g sm[33]=%g\n", v, cudaGetErrorString(e), ho[0], ho[1], ho[2]);
    return 0;
}
