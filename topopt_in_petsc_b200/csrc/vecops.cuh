// vecops.cuh -- declarations of the flat vector kernels / reductions in vecops.cu
#pragma once
#include "common.cuh"
namespace tob {
constexpr int kFlatThreads = 256;
int flat_grid(tob_ctx_impl* c, long long n);
// grid of a grid-stride kernel with a per-block prologue (flag waits, partial sums): never more blocks than are resident at once
// (a second wave would pay the prologue again and run at a fraction of the occupancy); per-kernel occupancy is queried once
int resident_grid(tob_ctx_impl* c, const void* kernel, int threads, long long n);
int finalize_sum(tob_ctx_impl* c, int nblocks, int nvals, int slot, bool allreduce);
int finalize_max(tob_ctx_impl* c, int nblocks, int nvals, int slot, bool allreduce);
int finalize_sum_to(tob_ctx_impl* c, int nblocks, int nvals, double* d_out, bool allreduce);
int finalize_max_to(tob_ctx_impl* c, int nblocks, int nvals, double* d_out, bool allreduce);
int pcg_update_xr(tob_ctx_impl* c, const LevelGrid& g, double* x, double* r, const double* p, const double* q,
                  const double* dinv, int s_num, int s_den, int slot_rr_rz);
int pcg_update_p(tob_ctx_impl* c, const LevelGrid& g, double* p, const double* z_or_r, const double* dinv, int s_num,
                 int s_den, int first);
int vec_axpby(tob_ctx_impl* c, double* y, const double* x, double a, double b, long long n);
int vec_pointwise_mult(tob_ctx_impl* c, double* y, const double* a, const double* b, long long n);
int vec_fill(tob_ctx_impl* c, double* y, double v, long long n);
int vec_restore_fixed(tob_ctx_impl* c, double* y, const double* u, const double* nf, long long n);
// single-rank Jacobi-PCG kernels that consume per-CTA partial sums directly (no finalize launches)
int pcg_update_xr_f(tob_ctx_impl* c, const LevelGrid& g, double* x, double* r, const double* p, const double* q, const float* dinv,
                    const double* pq_parts, int n_pq, const double* prev_parts, double* out_parts, int* nb_out, int xmode = 0,
                    double* ab = nullptr);
int pcg_update_p_f(tob_ctx_impl* c, const LevelGrid& g, double* p, const double* r, const float* dinv, const double* new_parts,
                   const double* old_parts, double* ab, int nparts);
int finalize_parts(tob_ctx_impl* c, const double* parts, int nb, int slot);
}  // namespace tob
