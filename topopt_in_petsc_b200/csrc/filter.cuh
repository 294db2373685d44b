// filter.cuh -- internal state of the Filter / PDEFilt replacement
#pragma once
#include "common.cuh"

namespace tob {

struct PdeFilterState;  // pdefilter.cu

struct tob_filter_impl {
    tob_ctx_impl* c = nullptr;
    int type = 1;
    double R = 0.0;
    // cone filter: element field with s ghost layers on each side (clipped by the domain)
    int s = 0;              // ElemConn
    bool tiled = true;      // TMA-tiled kernel for s = 1, 2 (else the direct-gather kernel)
    int gl0 = 0;            // global index of local layer 0
    int own0 = 0, own1 = 0; // owned local layers
    int nl = 0;             // stored local layers
    long long estride = 0;
    double* wtab = nullptr;   // (2s+1)^3 cone weights
    double* f_st = nullptr;   // filter-layout scratch field
    double* hs_st = nullptr;  // Hs = H 1 in filter layout
    double* tmp_nat = nullptr;
    double* tmp_nat2 = nullptr;
    std::vector<DevBuf> gbuf; // staging for host gradient vectors
    PdeFilterState* pde = nullptr;
    tob_solve_stats stats{};
};

int filter_create(tob_ctx_impl* c, int type, double rmin, tob_filter_impl** out);
void filter_destroy(tob_filter_impl* h);

// PDE (Helmholtz) filter, pdefilter.cu
int pdefilter_create(tob_filter_impl* h);
int pdefilter_apply(tob_filter_impl* h, const double* x_nat, double* out_nat);
void pdefilter_destroy(tob_filter_impl* h);

}  // namespace tob
