// elast.cuh -- internal state of the LinearElasticity replacement
#pragma once
#include "common.cuh"
#include "hex8.cuh"
#include "cgsr.cuh"

namespace tob {

struct GmgHierarchy;  // gmg.cu

struct tob_elast_impl {
    tob_ctx_impl* c = nullptr;
    const LevelGrid* g = nullptr;
    tob_solver_opts opts{};
    double nu = 0.3;
    Hex8Operator op;
    // stored-layout nodal vectors
    double *U = nullptr, *B = nullptr, *RHS = nullptr, *R = nullptr, *P = nullptr, *Q = nullptr, *Z = nullptr;
    double *DINV = nullptr, *NFREE = nullptr;
    float* DINV32 = nullptr;  // the Jacobi preconditioner of the one-rank PCG kernels, single precision (vecops.cu)
    unsigned char* mask = nullptr;
    // stored element fields
    double *xphys_st = nullptr, *E_st = nullptr;
    // natural-order scratch (device)
    double *nat_dof = nullptr, *nat_dof2 = nullptr, *nat_el = nullptr, *nat_el2 = nullptr;
    double Emin = 1e-9, Emax = 1.0, penal = 3.0;
    double bnorm = 0.0;
    bool have_density = false;
    tob_solve_stats stats{};
    GmgHierarchy* mg = nullptr;
    // optional per-launch SpMV timing (bench.py roofline): event pairs recorded around each launch
    bool profile = false;
    int prof_stride = 1;
    // CUDA graph of a block of multi-GPU Jacobi-PCG iterations (elast.cu)
    cudaGraphExec_t pcg_graph = nullptr;
    int pcg_graph_its = 0;
    long long pcg_graph_launches = 0;
    std::vector<cudaEvent_t> prof_ev;
    int prof_used = 0;
    // fp32 scratch of the tensor-core operator (hex8_tc32.cu)
    float *u32 = nullptr, *y32 = nullptr, *keT32 = nullptr;
    float4 *x4 = nullptr, *y4 = nullptr;
    float *E32 = nullptr, *tcB = nullptr, *tcB3 = nullptr;
    bool E32_valid = false;
    // single-reduction Jacobi-PCG (cgsr.cuh): u = D^-1 r lives in a peer-mapped buffer (the neighbours store its ghost planes)
    PeerVec sr_u;
    CgDev* sr_cg = nullptr;
    double* sr_upart = nullptr;
    CgPeers sr_pe;
    int sr_nbU = 0, sr_par = 0;
    cudaGraphExec_t sr_graph = nullptr;
    int sr_graph_its = 0;
    long long sr_graph_launches = 0;
};
int elast_profile_collect(tob_elast_impl* h);
int tc2_prepare(tob_elast_impl* h);  // buffers / operand blocks / fp32 modulus of the tensor-core operator (elast.cu)

int elast_create(tob_ctx_impl* c, double nu, const tob_solver_opts* opts, tob_elast_impl** out);
void elast_destroy(tob_elast_impl* h);
int elast_set_bc_load(tob_elast_impl* h, const unsigned char* fixed, const double* rhs);
int elast_set_density(tob_elast_impl* h, const double* xphys_nat_dev, double Emin, double Emax, double penal);
int elast_apply(tob_elast_impl* h, const double* u, double* y, int epi, bool mask_in, int kernel, const double* a0,
                int* nparts);
int elast_pcg(tob_elast_impl* h);
int mask_to_free(tob_ctx_impl* c, const LevelGrid& g, const unsigned char* mask, double* nf);

// geometric multigrid preconditioner (gmg.cu)
int gmg_setup(tob_elast_impl* h);
int gmg_update(tob_elast_impl* h);   // refresh coarse operators after a density change
int gmg_refresh_masks(tob_elast_impl* h);  // inject the Dirichlet mask to the coarse levels
int gmg_vcycle(tob_elast_impl* h, const double* r, double* z);
void gmg_destroy(tob_elast_impl* h);
void gmg_drop_graph(tob_elast_impl* h);  // after a change of solver options

}  // namespace tob
