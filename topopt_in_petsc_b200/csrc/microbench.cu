// microbench.cu -- FP64 pipe micro-benchmarks that size the hex8 kernel design:
// is a B200's DMMA (mma.sync.m8n8k4.f64) throughput additive to the vector DFMA pipe or shared with it?
// Results are committed under profiles/ and quoted in DESIGN.md.
#include "common.cuh"

namespace tob {

constexpr int MB_ITERS = 4096;

__global__ void __launch_bounds__(256) k_mb_dfma(double* out, double a, double b) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_mb_dmma(double* out, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mb_both(double* out, double a, double b) {
    double c[4][2];
    double r[8];
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            dmma884(c[i][0], c[i][1], a, b);
            r[2 * i] = fma(r[2 * i], a, b);
            r[2 * i + 1] = fma(r[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mb_copy(const double2* __restrict__ in, double2* __restrict__ out, long long n2) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace tob

using namespace tob;

extern "C" {

int tob_microbench_fp64(tob_ctx* ctx, int kind, double* tflops) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(ctx);
    double* d = c->d_partials;
    const int blocks = c->num_sms * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        TOB_CUDA(cudaEventRecord(c->ev0, c->stream));
        if (kind == 0) k_mb_dfma<<<blocks, threads, 0, c->stream>>>(d, 1.0000001, 1e-9);
        else if (kind == 1) k_mb_dmma<<<blocks, threads, 0, c->stream>>>(d, 1.0000001, 1e-9);
        else k_mb_both<<<blocks, threads, 0, c->stream>>>(d, 1.0000001, 1e-9);
        c->launches++;
        TOB_KERNEL_CHECK();
        TOB_CUDA(cudaEventRecord(c->ev1, c->stream));
        TOB_CUDA(cudaEventSynchronize(c->ev1));
        float ms;
        TOB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double nthreads = (double)blocks * threads;
    double flops;
    const double dfma = nthreads * MB_ITERS * 8.0 * 2.0;                    // 8 FMA per iteration per thread
    const double dmma = (nthreads / 32.0) * MB_ITERS * 4.0 * (8 * 8 * 4 * 2.0);  // 4 m8n8k4 per iteration per warp
    flops = kind == 0 ? dfma : (kind == 1 ? dmma : dfma + dmma);
    *tflops = flops / (best * 1e-3) / 1e12;
    return 0;
}

int tob_microbench_copy(tob_ctx* ctx, long long nbytes, double* gbs) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(ctx);
    void *a = nullptr, *b = nullptr;
    TOB_CUDA(cudaMalloc(&a, (size_t)nbytes));
    TOB_CUDA(cudaMalloc(&b, (size_t)nbytes));
    TOB_CUDA(cudaMemsetAsync(a, 1, (size_t)nbytes, c->stream));
    const long long n2 = nbytes / 16;
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        TOB_CUDA(cudaEventRecord(c->ev0, c->stream));
        k_mb_copy<<<c->num_sms * 16, 256, 0, c->stream>>>((const double2*)a, (double2*)b, n2);
        c->launches++;
        TOB_KERNEL_CHECK();
        TOB_CUDA(cudaEventRecord(c->ev1, c->stream));
        TOB_CUDA(cudaEventSynchronize(c->ev1));
        float ms;
        TOB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(a);
    cudaFree(b);
    *gbs = 2.0 * (double)nbytes / (best * 1e-3) / 1e9;
    return 0;
}

}  // extern "C"
