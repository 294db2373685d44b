// cgsr.cuh -- single-reduction (Chronopoulos-Gear) Jacobi-PCG: the multi-GPU Krylov iteration as TWO launches.
//
// Classical PCG needs two global reductions per iteration (p.Ap, then r.z), each a cross-GPU round trip on the critical
// path; with the halo exchange that was 6 launches and 3 flag round trips per iteration (round 1: 65 us per iteration on
// 8 GPUs for 25 us of kernel time).  The Chronopoulos-Gear recurrence carries s = A p along, so both inner products of an
// iteration are available after ONE operator application:
//     U:  [alpha, beta from last iteration's sums]   p = u + beta p ;  s = w + beta s ;  x += alpha p ;  r -= alpha s ;
//         u = D^-1 r  (also stored straight into the neighbours' ghost planes over NVLink) ;  partials of r.u and r.r
//         (r itself is not stored: u -= alpha D^-1 s and r = D u where the sums need it -- two vector streams fewer)
//     S:  w = A u  (z-march SpMV; only the CTAs that touch a ghost plane wait for the neighbour's flag) ; partials of w.u ;
//         the LAST CTA to finish sums all partials, stores the three sums into every peer's arena and raises a flag.
// The next U spins on the peers' flags and adds the per-rank sums in rank order (bitwise identical on all ranks).
// No halo kernel, no finalize kernels: 2 launches and 2 NVLink latencies per iteration.
#pragma once
#include "common.cuh"

namespace tob {

struct CgState {  // ping-pong: U reads st[par], writes st[par ^ 1]
    double gam_prev, alpha_prev, first;
    unsigned long long epoch;  // number of U launches so far on this context (monotone: the cross-GPU flag value)
};
struct CgDev {
    CgState st[2];
    double sums[3];         // single rank: r.u, r.r, w.u of the last (U, S) pair
    unsigned int ucounter;  // tickets: the boundary blocks of U (the last one raises the neighbours' flags), the CTAs of S (reduction)
    unsigned int scounter;
    double out[4];          // for the host: r.u, r.r, w.u, alpha
};

// u64 indices in the arena flag page used by this path (comm.cu owns indices < 40)
enum { F_CG_FROM_LO = 40, F_CG_FROM_HI = 41, F_CG_RED0 = 64 };
constexpr int kCgSlot = 120;  // doubles 120..122 of a reduction slot carry the three sums

struct CgPeers {  // per-launch description of the neighbourhood (by value in the kernel parameters)
    int n = 1, rank = 0;
    unsigned char* arena = nullptr;         // my arena (flags + reduction slots)
    unsigned char* peer[kMaxRanks] = {};    // every rank's arena
    double* u_lo_ghost = nullptr;           // lower neighbour's TOP ghost plane of u (null if none)
    double* u_hi_ghost = nullptr;           // upper neighbour's BOTTOM ghost plane of u
    // TOB_CG_TRACE=1: %globaltimer stamps, 8 per epoch (slot = epoch % trace_cap):
    //   0 U entry, 1 U sums arrived, 2 U last block done (flags raised), 3 S entry, 4 S ghost planes arrived, 5 S last CTA published
    unsigned long long* trace = nullptr;
    int trace_cap = 0;
};
// Programmatic dependent launch (TOB_PDL=1, off by default) for the U/S pair: the next kernel of the pair is made resident while
// the current one drains and parks at pdl_wait().  Measured on 2 GPUs (cfg2): the S -> U gap shrinks 2.2 -> 1.7 us but the U -> S
// hand-over does not (an S CTA needs a whole SM, so it cannot start before the U blocks of that SM are gone) and the iteration is
// 1-2 % slower -- kept as a switch, not as the default.  Both instructions are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
    static const bool on = getenv("TOB_PDL") && atoi(getenv("TOB_PDL")) != 0;
    return on;
}

// System-scope release pattern with ONE fence: fence.acq_rel.sys followed by relaxed system-scope stores (PTX memory model: a
// fence + relaxed store is a release pattern).  The first version used __threadfence_system() (MEMBAR.SC.SYS) and then a
// st.release.sys per flag (each a MEMBAR.ALL.SYS + store): 3-4 system fences in series on the critical path of every iteration.
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void cg_trace(const CgPeers& pe, unsigned long long epoch, int k) {
    if (pe.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        pe.trace[(epoch % (unsigned long long)pe.trace_cap) * 8 + k] = t;
    }
}

size_t cg_red_off(int n, int parity, int rank);  // byte offset of a reduction slot inside an arena (comm.cu)

// peer-mapped vector: cudaMalloc + IPC handle exchange (collective).  peer[k] = rank k's buffer in this address space.
struct PeerVec {
    double* mine = nullptr;
    double* peer[kMaxRanks] = {};
    size_t bytes = 0;
};
int peer_vec_alloc(tob_ctx_impl* c, size_t nbytes, PeerVec* out);
void peer_vec_free(tob_ctx_impl* c, PeerVec* v);

// U kernel.  mode 0 = set-up (u = dinv r only, first = 1), mode 1 = iteration.  par = parity of the state it reads.
// r is read by the set-up launch only (the iteration carries the residual in u = D^-1 r); dinv is the fp32 copy of the inverse diagonal
int cg_update(tob_ctx_impl* c, const LevelGrid& g, CgDev* cg, int par, int mode, double* x, const double* r, double* u, const double* w,
              double* p, double* s, const float* dinv, double* upart, int* nbU, const CgPeers& peers, int xmode = 0);

}  // namespace tob
