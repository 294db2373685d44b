// hex8.cuh -- matrix-free 8-node-hex stiffness operator (replaces the assembled AIJ matrix of
// LinearElasticity::AssembleStiffnessMatrix, LinearElasticity.cc:487-549, and MatMult on it).
#pragma once
#include "common.cuh"

namespace tob {

// KE block-diagonalised by the reflection symmetries of a rectangular hex:
// with T = W (x) I3, W the 8x8 Walsh-Hadamard matrix over the node offsets (a,b,g),
// D = T KE T^T / 64 decomposes into 8 blocks of 3x3, block rho coupling the three basis
// vectors (mode rho^(1<<c), component c), c = 0..2.   KE = T^T D T.
struct Hex8Blocks {
    double b[8][9];  // b[rho][3*c + c2]
};

struct Hex8Operator {
    double KE[576];      // reference ordering, row-major ke[j+24*i] (LinearElasticity.cc:841-998)
    Hex8Blocks blocks;
    double kdiag[3];     // KE[(n,c),(n,c)], identical for the 8 nodes of a box element
    double* d_KE = nullptr;  // device copy for the naive cross-check kernel
    double offblock_max = 0; // largest entry of D outside the 8 blocks (structure check)
};

void hex8_element_matrix(double dx, double dy, double dz, double nu, double* ke576);
int hex8_build_operator(double dx, double dy, double dz, double nu, Hex8Operator& op);
void hex8_free_operator(Hex8Operator& op);

enum SpmvEpilogue { EPI_Y = 0, EPI_DOT = 1, EPI_RESID = 2, EPI_CHEB = 3, EPI_CGDOT = 4 };
struct CgDev;
struct CgPeers;
enum SpmvKernel { SPMV_AUTO = 0, SPMV_NAIVE = 1, SPMV_ZM8 = 2, SPMV_ZM16 = 3, SPMV_ZM12 = 4, SPMV_ZM8R = 5, SPMV_ZM8RF = 6, SPMV_ZM16RF = 7, SPMV_ZM12RF = 8, SPMV_ZM16R = 9, SPMV_ZM16U4 = 10, SPMV_ZM16M4 = 11 };

struct SpmvArgs {
    const double* u = nullptr;           // stored nodal vector (ghost planes already exchanged)
    double* y = nullptr;                 // stored nodal vector (owned planes written)
    const double* E = nullptr;           // stored element modulus field
    const unsigned char* mask = nullptr; // stored Dirichlet mask
    int epi = EPI_Y;
    bool mask_in = true;                 // input still needs N applied (naive kernel only; z-march callers pre-mask)
    int kernel = SPMV_AUTO;
    // epilogue operands (stored layout)
    const double* a0 = nullptr;          // RESID/CHEB: b
    const double* a1 = nullptr;          // (unused by the z-march kernels)
    const float* a1f = nullptr;          // CHEB: dinv, fp32 copy (the smoother is a preconditioner: half a vector stream less per step)
    double* o1 = nullptr;                // CHEB: d (in/out)
    double s0 = 0, s1 = 0;               // CHEB: d = s0*d + s1*dinv*(b - A u); y = u + d
    int l2hint = 0;                      // L2 eviction hints for the TMA loads (Jacobi-PCG; tob_ctx_impl::l2hint)
    // EPI_CGDOT (cgsr.cuh): y = A u, partials of u.y, last CTA reduces U's and its own partials and publishes the sums
    CgDev* cg = nullptr;
    int cgpar = 0;
    const double* upart = nullptr;
    int nbU = 0;
    const CgPeers* pe = nullptr;
};

// y = N.(K (N.u)) + (1-N).u on the owned planes of level grid g; *nparts = number of per-CTA partial
// sums written to ctx->d_partials for EPI_DOT / EPI_RESID (0 otherwise).
int hex8_spmv(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const SpmvArgs& a, int* nparts);

// single-precision operator on the tensor cores (hex8_tc32.cu: tcgen05.mma kind::tf32, accumulators in tensor memory)
int hex8_spmv_tc32(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const float* u32, float* y32, const double* E,
                   const unsigned char* mask, float** keT_cache);
// generation 2 (implicit im2col through overlapping UMMA descriptors) on node-interleaved fp32 vectors ("V4": float4 per node)
int tc2_build_B(const Hex8Operator& op, float** dB);
int hex8_tc32b(tob_ctx_impl* c, const LevelGrid& g, const float* dB, const float4* x4, float4* y4, const float* E32, const unsigned char* mask,
               int epi, const float4* b4, const float4* dinv4, float4* d4, float s0, float s1);
int tc3_build_B(const Hex8Operator& op, float** dB);
int hex8_tc32c(tob_ctx_impl* c, const LevelGrid& g, const float* dB, const float4* x4, float4* y4, const float* E32, const unsigned char* mask,
               int epi, const float4* b4, const float4* dinv4, float4* d4, float s0, float s1);
long long v4_plane_nodes(const LevelGrid& g);  // nodes per plane of a V4 vector (allocation: v4_plane_nodes * nzl float4)
int vec_st_to_v4(tob_ctx_impl* c, const LevelGrid& g, const double* st, float4* v4);
int vec_v4_to_st(tob_ctx_impl* c, const LevelGrid& g, const float4* v4, double* st);
int vec_d2f(tob_ctx_impl* c, const double* a, float* b, long long n);
int vec_f2d(tob_ctx_impl* c, const float* a, double* b, long long n);

// dinv = 1/diag(A) (1 at clamped dofs), stored layout
int hex8_diag_inv(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const double* E,
                  const unsigned char* mask, double* dinv, bool invert);

// per-element u^T KE u, objective partial sums and sensitivities (LinearElasticity.cc:403-437).
// xphys/E in stored layout, dfdx in NATURAL owned order. d_scalars[slot] = sum E uKu, [slot+1] = sum xphys.
int hex8_compliance(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const double* u, const double* xphys,
                    const double* E, double penal, double Emin, double Emax, double* dfdx_nat, int slot);

// E = Emin + x^penal (Emax-Emin) on a stored element field (LinearElasticity.cc:519)
int simp_modulus(tob_ctx_impl* c, const double* x, double* E, long long n, double Emin, double Emax, double penal);

}  // namespace tob
