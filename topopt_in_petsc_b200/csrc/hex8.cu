// hex8.cu -- matrix-free hex8 stiffness operator for sm_100a.
//
// Hot kernel: k_hex8_zmarch.  One CTA owns an (x,y) tile of node columns and marches along z.
// Every z-step a TMA bulk-tensor load (cp.async.bulk.tensor, mbarrier complete_tx) brings the next
// nodal plane tile (3 components) and the element-modulus tile of the layer below it into a
// 4-stage shared-memory ring (the byte-wide Dirichlet mask of the own node comes through a plain
// load issued one step ahead).  Each thread owns one element column:
// it forms the Walsh-Hadamard "face modes" of the new plane, combines them with the face modes
// of the previous plane kept in registers, applies the 8 3x3 blocks of the block-diagonalised
// element matrix scaled by the SIMP modulus, and transforms back stage by stage; partial results
// are shared between the elements around a face / edge / node (registers along z, one shared-
// memory exchange along y, one warp shuffle along x), so each nodal value is loaded once and
// written once and the FP64 work is ~180 DP ops per element instead of 576 FMA.
//
// The dense cross-check kernel k_hex8_naive applies the 24x24 KE rows per node directly.
#include "hex8.cuh"
#include "tma.cuh"
#include "vecops.cuh"
#include "cgsr.cuh"

namespace tob {

// ---------------------------------------------------------------------------------------
// host: element matrix and its symmetry-adapted block form
// ---------------------------------------------------------------------------------------
static const int kSgn[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                               {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};  // LinearElasticity.cc:118-120 node order
// binary node index q = a + 2b + 4g  ->  reference local node
static const int kBin2Ref[8] = {0, 1, 3, 2, 4, 5, 7, 6};

void hex8_element_matrix(double dx, double dy, double dz, double nu, double* ke) {
    // 2x2x2 Gauss quadrature of B^T C B on the box element (LinearElasticity.cc:841-998; E = 1).
    // For an axis-aligned box the Jacobian is diag(dx,dy,dz)/2, so physical shape-function
    // gradients are the parent gradients scaled by 2/d.
    const double lam = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));  // :887
    const double mu = 1.0 / (2.0 * (1.0 + nu));               // :888
    const double gp = 0.577350269189626;                      // :898
    const double detJ = dx * dy * dz / 8.0;
    for (int i = 0; i < 576; i++) ke[i] = 0.0;
    for (int a = 0; a < 8; a++) {
        const double xi = (a & 1) ? gp : -gp, eta = (a & 2) ? gp : -gp, zeta = (a & 4) ? gp : -gp;
        double G[8][3];  // physical gradient of N_n
        for (int n = 0; n < 8; n++) {
            const double sx = kSgn[n][0], sy = kSgn[n][1], sz = kSgn[n][2];
            G[n][0] = 0.125 * sx * (1 + sy * eta) * (1 + sz * zeta) * (2.0 / dx);
            G[n][1] = 0.125 * sy * (1 + sx * xi) * (1 + sz * zeta) * (2.0 / dy);
            G[n][2] = 0.125 * sz * (1 + sx * xi) * (1 + sy * eta) * (2.0 / dz);
        }
        // ke[(n,c),(m,d)] += detJ * [ lam G_nc G_md + mu G_nd G_mc + mu delta_cd (G_n . G_m) ]
        for (int n = 0; n < 8; n++)
            for (int m = 0; m < 8; m++) {
                const double gg = G[n][0] * G[m][0] + G[n][1] * G[m][1] + G[n][2] * G[m][2];
                for (int cc = 0; cc < 3; cc++)
                    for (int d = 0; d < 3; d++) {
                        double v = lam * G[n][cc] * G[m][d] + mu * G[n][d] * G[m][cc];
                        if (cc == d) v += mu * gg;
                        ke[(3 * n + cc) * 24 + 3 * m + d] += detJ * v;
                    }
            }
    }
}

int hex8_build_operator(double dx, double dy, double dz, double nu, Hex8Operator& op) {
    hex8_element_matrix(dx, dy, dz, nu, op.KE);
    // D = T KEbin T^T / 64
    static double Kb[24][24], TK[24][24], D[24][24];
    for (int q = 0; q < 8; q++)
        for (int c = 0; c < 3; c++)
            for (int q2 = 0; q2 < 8; q2++)
                for (int c2 = 0; c2 < 3; c2++)
                    Kb[3 * q + c][3 * q2 + c2] = op.KE[(3 * kBin2Ref[q] + c) * 24 + 3 * kBin2Ref[q2] + c2];
    auto W = [](int m, int q) { return (__builtin_popcount(m & q) & 1) ? -1.0 : 1.0; };
    for (int m = 0; m < 8; m++)
        for (int c = 0; c < 3; c++)
            for (int col = 0; col < 24; col++) {
                double s = 0;
                for (int q = 0; q < 8; q++) s += W(m, q) * Kb[3 * q + c][col];
                TK[3 * m + c][col] = s;
            }
    for (int row = 0; row < 24; row++)
        for (int m = 0; m < 8; m++)
            for (int c = 0; c < 3; c++) {
                double s = 0;
                for (int q = 0; q < 8; q++) s += TK[row][3 * q + c] * W(m, q);
                D[row][3 * m + c] = s / 64.0;
            }
    double dmax = 0, off = 0;
    for (int m = 0; m < 8; m++)
        for (int c = 0; c < 3; c++)
            for (int m2 = 0; m2 < 8; m2++)
                for (int c2 = 0; c2 < 3; c2++) {
                    const double v = fabs(D[3 * m + c][3 * m2 + c2]);
                    dmax = fmax(dmax, v);
                    if ((m ^ (1 << c)) != (m2 ^ (1 << c2))) off = fmax(off, v);
                }
    op.offblock_max = off;
    if (off > 1e-12 * dmax) TOB_FAIL("element matrix is not block-diagonal in the reflection-symmetry basis");
    for (int rho = 0; rho < 8; rho++)
        for (int c = 0; c < 3; c++)
            for (int c2 = 0; c2 < 3; c2++)
                op.blocks.b[rho][3 * c + c2] = D[3 * (rho ^ (1 << c)) + c][3 * (rho ^ (1 << c2)) + c2];
    // rigid translations (mode 0, component c) lie in block rho = 1<<c: that row/column must vanish
    for (int c = 0; c < 3; c++)
        for (int c2 = 0; c2 < 3; c2++) {
            if (fabs(op.blocks.b[1 << c][3 * c + c2]) > 1e-12 * dmax || fabs(op.blocks.b[1 << c][3 * c2 + c]) > 1e-12 * dmax)
                TOB_FAIL("translation mode is not in the null space of KE");
            op.blocks.b[1 << c][3 * c + c2] = 0.0;
            op.blocks.b[1 << c][3 * c2 + c] = 0.0;
        }
    // mixed blocks: entry (c,d), c != d, vanishes unless bits c and d of rho agree (the gen-2 kernel skips those products)
    for (int rho = 0; rho < 8; rho++)
        for (int c = 0; c < 3; c++)
            for (int c2 = 0; c2 < 3; c2++)
                if (c != c2 && (((rho >> c) & 1) != ((rho >> c2) & 1))) {
                    if (fabs(op.blocks.b[rho][3 * c + c2]) > 1e-12 * dmax) TOB_FAIL("symmetry block has an unexpected coupling");
                    op.blocks.b[rho][3 * c + c2] = 0.0;
                }
    for (int c = 0; c < 3; c++) op.kdiag[c] = op.KE[c * 24 + c];
    for (int n = 1; n < 8; n++)
        for (int c = 0; c < 3; c++)
            if (fabs(op.KE[(3 * n + c) * 24 + 3 * n + c] - op.kdiag[c]) > 1e-12 * dmax) TOB_FAIL("KE diagonal not node-invariant");
    TOB_CUDA(cudaMalloc(&op.d_KE, sizeof(double) * 576));
    TOB_CUDA(cudaMemcpy(op.d_KE, op.KE, sizeof(double) * 576, cudaMemcpyHostToDevice));
    return 0;
}

void hex8_free_operator(Hex8Operator& op) {
    if (op.d_KE) cudaFree(op.d_KE);
    op.d_KE = nullptr;
}

// ---------------------------------------------------------------------------------------
// naive dense cross-check kernel: one thread per owned node, 8 x (3 x 24) KE rows
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_hex8_naive(const double* __restrict__ u, double* __restrict__ y, const double* __restrict__ E,
             const unsigned char* __restrict__ mask, const double* __restrict__ KE, int nxn, int nyn, int nzl, int px,
             int pe, int pm, long long nstride, long long pstride, long long estride, long long mstride, int z0, int z1,
             int mask_in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int z = z0 + blockIdx.z;
    if (i >= nxn || z >= z1) return;
    const int nex = nxn - 1, ney = nyn - 1, nzle = nzl - 1;
    double acc[3] = {0, 0, 0};
    for (int dk = -1; dk <= 0; dk++)
        for (int dj = -1; dj <= 0; dj++)
            for (int di = -1; di <= 0; di++) {
                const int ei = i + di, ej = j + dj, ek = z + dk;
                if (ei < 0 || ei >= nex || ej < 0 || ej >= ney || ek < 0 || ek >= nzle) continue;
                const double Ee = E[ek * estride + (long long)ej * pe + ei];
                const int q = (-di) + 2 * (-dj) + 4 * (-dk);
                const int rloc = (q == 2) ? 3 : (q == 3) ? 2 : (q == 6) ? 7 : (q == 7) ? 6 : q;
                double s[3] = {0, 0, 0};
                for (int n = 0; n < 8; n++) {
                    const int a = (n == 1 || n == 2 || n == 5 || n == 6), b = (n == 2 || n == 3 || n == 6 || n == 7), g = n >> 2;
                    const int ni = ei + a, nj = ej + b, nk = ek + g;
                    const unsigned char m = mask_in ? mask[nk * mstride + (long long)nj * pm + ni] : 0;
                    for (int c2 = 0; c2 < 3; c2++) {
                        double uv = u[nk * pstride + c2 * nstride + (long long)nj * px + ni];
                        if (m & (1 << c2)) uv = 0.0;
                        for (int c = 0; c < 3; c++) s[c] = fma(__ldg(&KE[(3 * rloc + c) * 24 + 3 * n + c2]), uv, s[c]);
                    }
                }
                for (int c = 0; c < 3; c++) acc[c] = fma(Ee, s[c], acc[c]);
            }
    const unsigned char m = mask[z * mstride + (long long)j * pm + i];
    for (int c = 0; c < 3; c++) {
        const long long idx = z * pstride + c * nstride + (long long)j * px + i;
        y[idx] = (m & (1 << c)) ? u[idx] : acc[c];
    }
}

// ---------------------------------------------------------------------------------------
// z-march kernel
// ---------------------------------------------------------------------------------------

struct ZmParams {
    int nxn, nyn, nzl;
    int px, pm;
    long long mstride;
    const unsigned char* mask;
    long long nstride, pstride;
    int zout0, zout1;  // local node planes written: [zout0, zout1)
    int zchunk;        // planes per CTA along z
    double* y;
    const double* a0;
    const double* a1;
    const float* a1f;
    double* o1;
    double s0, s1;
    double* partials;
    int l2hint;  // Jacobi-PCG only (common.cuh): 1 = modulus tiles evict_first, 2 = input vector (p) evict_last
    // EPI_CGDOT (single-reduction PCG, cgsr.cuh): wait for the neighbours' ghost planes, last CTA reduces and publishes
    CgDev* cg;
    int cgpar;            // parity of the state the preceding U kernel wrote (holds the epoch of this operand)
    const double* upart;  // U's per-CTA partials: [0,nbU) r.u, [nbU,2nbU) r.r
    int nbU;
    CgPeers pe;
    size_t red_off0, red_off1;
};

// TMA tiles are ZM_W = 34 doubles wide: with a non-swizzled tensor map the innermost start
// coordinate must be 16-byte aligned (measured: odd fp64 origins raise "illegal instruction"), so the
// box starts at the even column bx*30-2 and thread tx reads column tx+1.
constexpr int ZM_W = 34;
template <int TY, bool FLAGS, int ROWS = TY>
struct ZmSmem {
    static constexpr int NT = 32 * TY;
    static constexpr int NC = ZM_W * ROWS;        // doubles per component tile (ROWS node rows in the TMA box)
    static constexpr int U_BYTES = (3 * NC * 8 + 127) / 128 * 128;  // TMA destinations must be 128-byte aligned
    static constexpr int E_BYTES = (NC * 8 + 127) / 128 * 128;
    static constexpr int TX_BYTES = 4 * NC * 8;                     // bytes the two TMA loads deliver per stage
    static constexpr int STAGE_BYTES = U_BYTES + E_BYTES;
    // TMA ring depth (a plane is consumed by exactly one step, so the prefetch distance is STAGES-1 planes)
    static constexpr int STAGES = FLAGS ? 4 : 6;
    // exchange-buffer ring (y-neighbour hand-off).  FLAGS: no reverse (reader -> writer) signal exists; a row can
    // start step t only after every row finished step t - STAGES (that is when the producer may refill the slot
    // of plane t + 1), so a ring of STAGES exchange slots can never be overwritten before it was read.
    static constexpr int XCH_DEPTH = FLAGS ? STAGES : 2;
    static constexpr int XCH_BYTES = XCH_DEPTH * 6 * NT * 8;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES + XCH_BYTES;
    // mbarriers: full[STAGES]; FLAGS adds empty[STAGES] and xfull[TY][XCH_DEPTH]
    static constexpr int NBARS = STAGES + (FLAGS ? STAGES + TY * XCH_DEPTH : 0);
    static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 128;
};

// the 8 symmetric 3x3 blocks as 16-byte pairs (b00 b01 | b02 b11 | b12 b22) so that one LDCU.128 fetches two
struct alignas(16) Hex8BlocksPacked {
    double2 k[8][3];
};

// One CTA = 32 x TY node columns marching along z; the 24 doubles carried from one z-step to the next (lower-face
// modes + upper-face contributions) stay in registers.
// FLAGS = false: one __syncthreads per z-step orders the y-neighbour exchange; thread 0 refills the TMA ring.
// FLAGS = true : no CTA-wide barrier inside the march.  Warp row ty hands its y-exchange data to row ty+1 through one
// mbarrier per (row, exchange slot); the top row (which has no elements of its own) is the TMA producer and
// recycles a ring slot when all TY-1 consumer rows have arrived on its "empty" mbarrier.  Waiting warps sit in
// mbarrier.try_wait (hardware-suspended), not in a shared-memory spin loop -- the earlier spin-flag version of this
// idea ran 2.4x slower than the barrier kernel (profiles/r01_ncu_zmarch_ty8_flags_256.txt).
template <int TY, int EPI, bool FLAGS>
__global__ void __launch_bounds__(32 * TY, (TY <= 8) ? 2 : 1)
k_hex8_zmarch(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmE,
              const __grid_constant__ Hex8BlocksPacked KB, const __grid_constant__ ZmParams P) {
    using L = ZmSmem<TY, FLAGS>;
    constexpr int NT = L::NT;
    constexpr int NC = L::NC;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double red[32];
    double* xch = reinterpret_cast<double*>(smem_raw + L::STAGES * L::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::BAR_OFF);  // full[STAGES]
    uint64_t* ebars = bars + L::STAGES;                                   // FLAGS: empty[STAGES]
    uint64_t* xbars = bars + 2 * L::STAGES;                               // FLAGS: xfull[TY][XCH_DEPTH]

    const int tid = threadIdx.x;
    const int tx = tid & 31;
    const int ty = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp row; the broadcast lets ptxas treat it as warp-uniform
    const int i0 = blockIdx.x * 30 - 1, j0 = blockIdx.y * (TY - 2) - 1;
    const int i = i0 + tx, j = j0 + ty;
    const int z0 = P.zout0 + blockIdx.z * P.zchunk;
    const int z1 = min(z0 + P.zchunk, P.zout1);
    // rows 0..TY-2 own an element row each (lane 31 computes a throw-away column: no divergence inside the step);
    // row TY-1 only closes the tile: idle in the barrier variant, TMA producer in the mbarrier variant
    const bool rowcompute = ty < TY - 1;
    const bool outthr = (tx >= 1) && (tx <= 30) && (ty >= 1) && (ty <= TY - 2) && (i < P.nxn) && (j < P.nyn);
    const int pfirst = z0 - 1;        // first node plane consumed (prologue)
    const int nplanes = z1 - z0 + 2;  // planes pfirst .. z1
    double part = 0.0;

    if (z0 < z1) {
        if (tid == 0) {
            for (int s = 0; s < L::STAGES; s++) mbar_init(&bars[s], 1);
            if (FLAGS) {
                for (int s = 0; s < L::STAGES; s++) mbar_init(&ebars[s], TY - 1);
                for (int s = 0; s < TY * L::XCH_DEPTH; s++) mbar_init(&xbars[s], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();

        auto issue = [&](int pidx) {  // pidx = plane counter 0..nplanes-1, plane = pfirst + pidx
            const int s = pidx % L::STAGES;
            unsigned char* st = smem_raw + s * L::STAGE_BYTES;
            const int p = pfirst + pidx;
            mbar_expect_tx(&bars[s], L::TX_BYTES);
            if (P.l2hint & 2) tma_load_4d_hint(st, &tmU, &bars[s], i0 - 1, j0, 0, p, l2_policy_last());
            else tma_load_4d(st, &tmU, &bars[s], i0 - 1, j0, 0, p);
            if (P.l2hint & 1) tma_load_3d_hint(st + L::U_BYTES, &tmE, &bars[s], i0 - 1, j0, p - 1, l2_policy_first());
            else tma_load_3d(st + L::U_BYTES, &tmE, &bars[s], i0 - 1, j0, p - 1);
        };
        if (FLAGS) {
            // producer row: plane q reuses the slot of plane q - STAGES, which every consumer row releases at the end
            // of the step that consumed it (step q - STAGES - 1; the prologue for plane 0)
            if (ty == TY - 1) {
                for (int q = 0; q < nplanes; q++) {  // whole warp waits (keeps the warp converged), lane 0 issues
                    if (q >= L::STAGES) mbar_wait(&ebars[q % L::STAGES], ((q / L::STAGES) - 1) & 1);
                    if (tx == 0) issue(q);
                }
            }
        } else if (tid == 0) {
            for (int q = 0; q < L::STAGES && q < nplanes; q++) issue(q);
        }

        // Dirichlet mask byte of my own node, one plane per step through a running pointer (0 outside the grid)
        const bool innode = (i >= 0) && (i < P.nxn) && (j >= 0) && (j < P.nyn);
        const unsigned char* mptr = P.mask + ((long long)pfirst * P.mstride + (long long)j * P.pm + i);
        const int o = ty * ZM_W + tx + 1;  // my node inside a component tile
        const int o2 = ty * 32 + tx;       // my slot in the exchange buffers
        long long yoff = (long long)pfirst * P.pstride + (long long)j * P.px + i;  // my node on the plane completed by the current step

        // face modes of one plane tile: X stage on rows j and j+1, then Y stage; also hands back my own nodal
        // value (the Dirichlet / dot-product operand of the NEXT step's output, kept in registers)
        auto face_modes = [&](const unsigned char* st, double (&fm)[12], double (&own)[3]) {
            const double* su = reinterpret_cast<const double*>(st);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double a = su[c * NC + o], b = su[c * NC + o + 1], cc = su[c * NC + o + ZM_W], d = su[c * NC + o + ZM_W + 1];
                const double s0 = a + b, d0 = a - b, s1 = cc + d, d1 = cc - d;
                fm[4 * c + 0] = s0 + s1;
                fm[4 * c + 1] = d0 + d1;
                fm[4 * c + 2] = s0 - s1;
                fm[4 * c + 3] = d0 - d1;
                own[c] = a;
            }
        };

        // carried state: A = plane below the current element layer, B = plane above (becomes A at the end of the step).
        // (Unrolling the march by two with A/B swapped would save the ~30 register moves per step, but ptxas then loses
        // the warp-uniformity of the constant loads in the mbarrier variant -- LDC into vector registers, 400 B of spills.)
        double fA[12], fB[12], upr[12], uA[3], uB[3];
        unsigned mA = 0, mB = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) { fA[k] = 0.0; fB[k] = 0.0; upr[k] = 0.0; }
#pragma unroll
        for (int c = 0; c < 3; c++) { uA[c] = 0.0; uB[c] = 0.0; }

        // ---- prologue: plane pfirst ------------------------------------------------------
        if (!FLAGS || rowcompute) {
            mbar_wait(&bars[0], 0);
            mA = (innode && pfirst >= 0) ? (unsigned)*mptr : 0u;
            mptr += P.mstride;
            if (rowcompute) face_modes(smem_raw, fA, uA);
            if (FLAGS) {
                __syncwarp();
                if (tx == 0) mbar_arrive(&ebars[0]);
            }
        }

        // ---- one z-step: element layer e = pfirst + t, upper plane pfirst + t + 1 ---------------
        // fl/uo/mo: lower-face modes, own nodal value and mask of plane e (from the previous step);
        // fn/un/mn: the same for plane e + 1 (for the next step)
        auto step = [&](const int t, const double (&fl)[12], double (&fn)[12], const double (&uo)[3], double (&un)[3],
                        const unsigned mo, unsigned& mn) {
            const int pidx = t + 1;
            const int s = pidx % L::STAGES;
            const unsigned char* st = smem_raw + s * L::STAGE_BYTES;
            mn = (innode && pfirst + pidx < P.nzl) ? (unsigned)*mptr : 0u;  // consumed one step later
            mptr += P.mstride;
            double* xw = xch + (t % L::XCH_DEPTH) * 6 * NT;
            if (!FLAGS && !rowcompute) {  // idle top row of the barrier variant
                mbar_wait(&bars[s], (pidx / L::STAGES) & 1);
                __syncthreads();
                return;
            }
            mbar_wait(&bars[s], (pidx / L::STAGES) & 1);
            // Block constants straight from the kernel-parameter constant bank as uniform-register operands of the
            // DP instructions (SASS: LDCU.128 c[0x0][UR+imm] ... DFMA R, R, UR, R).  The index carries a term that is
            // zero for every reachable t but opaque to ptxas: with a loop-invariant address it hoists all 48 loads
            // and then thrashes the 63 uniform registers (R2UR spills); re-loading per step keeps the shared-memory
            // pipe (24 of ~95 wavefronts per warp-step were these constants) and 12 vector registers free.
            const double2* kz = &KB.k[0][0] + (t >> 30);

            double ex0[3][2];  // edge (row j) contributions of my face
            {
                face_modes(st, fn, un);
                const double Ee = reinterpret_cast<const double*>(st + L::U_BYTES)[o];
                // Z stage: element modes  uh[c][ma + 2 mb + 4 mg]
                double uh[3][8];
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int f = 0; f < 4; f++) {
                        uh[c][f] = fl[4 * c + f] + fn[4 * c + f];
                        uh[c][4 + f] = fl[4 * c + f] - fn[4 * c + f];
                    }
                // 8 symmetric blocks of 3x3 (rho = 1,2,4 have a zero row/column)
                double yh[3][8];
#pragma unroll
                for (int rho = 0; rho < 8; rho++) {
                    const double in0 = uh[0][rho ^ 1], in1 = uh[1][rho ^ 2], in2 = uh[2][rho ^ 4];
                    const double2 k0 = kz[rho * 3 + 0], k1 = kz[rho * 3 + 1], k2 = kz[rho * 3 + 2];
                    const double b00 = k0.x, b01 = k0.y, b02 = k1.x, b11 = k1.y, b12 = k2.x, b22 = k2.y;
                    double o0, o1, o2;
                    if (rho == 1) {
                        o0 = 0.0;
                        o1 = b11 * in1 + b12 * in2;
                        o2 = b12 * in1 + b22 * in2;
                    } else if (rho == 2) {
                        o1 = 0.0;
                        o0 = b00 * in0 + b02 * in2;
                        o2 = b02 * in0 + b22 * in2;
                    } else if (rho == 4) {
                        o2 = 0.0;
                        o0 = b00 * in0 + b01 * in1;
                        o1 = b01 * in0 + b11 * in1;
                    } else {
                        o0 = b00 * in0 + b01 * in1 + b02 * in2;
                        o1 = b01 * in0 + b11 * in1 + b12 * in2;
                        o2 = b02 * in0 + b12 * in1 + b22 * in2;
                    }
                    yh[0][rho ^ 1] = o0;  // the modulus is applied in the Z' stage below (one FMA instead of mul + add)
                    yh[1][rho ^ 2] = o1;
                    yh[2][rho ^ 4] = o2;
                }
                // Z' stage + face totals (lower face of this element + upper face of the previous one), scaled by E
                double ht[3][4];
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int f = 0; f < 4; f++) {
                        const double sl = yh[c][f] + yh[c][4 + f], su2 = yh[c][f] - yh[c][4 + f];
                        ht[c][f] = fma(Ee, sl, upr[4 * c + f]);
                        upr[4 * c + f] = Ee * su2;
                    }
                // Y' stage: row j part stays, row j+1 part goes to the thread above
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int ma = 0; ma < 2; ma++) {
                        ex0[c][ma] = ht[c][ma] + ht[c][ma + 2];
                        xw[(2 * c + ma) * NT + o2] = ht[c][ma] - ht[c][ma + 2];
                    }
            }
            if (FLAGS) {
                __syncwarp();
                if (tx == 0 && ty < TY - 2) mbar_arrive(&xbars[ty * L::XCH_DEPTH + t % L::XCH_DEPTH]);
                if (ty >= 1) mbar_wait(&xbars[(ty - 1) * L::XCH_DEPTH + t % L::XCH_DEPTH], (t / L::XCH_DEPTH) & 1);
            } else {
                __syncthreads();
                // every thread is past plane t (consumed by step t-1 / the prologue): refill its ring slot
                if (tid == 0 && t + L::STAGES < nplanes) issue(t + L::STAGES);
            }

            // edge totals, X' stage, node totals
            double v[3], vr[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double e0 = ex0[c][0], e1 = ex0[c][1];
                if (ty >= 1) {
                    e0 += xw[(2 * c + 0) * NT + o2 - 32];
                    e1 += xw[(2 * c + 1) * NT + o2 - 32];
                }
                v[c] = e0 + e1;
                vr[c] = e0 - e1;
            }
#pragma unroll
            for (int c = 0; c < 3; c++) v[c] += __shfl_up_sync(0xffffffffu, vr[c], 1);

            const int e = pfirst + t;  // node plane completed by this step
            if (outthr && e >= z0) {
                const long long idx0 = yoff;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const long long idx = idx0 + c * P.nstride;
                    const bool fixed = (mo >> c) & 1u;
                    const double au = fixed ? uo[c] : v[c];
                    if (EPI == EPI_Y) {
                        P.y[idx] = au;
                    } else if (EPI == EPI_DOT) {
                        P.y[idx] = au;
                        part = fma(uo[c], au, part);
                    } else if (EPI == EPI_RESID) {
                        const double r = P.a0[idx] - au;
                        P.y[idx] = r;
                        part = fma(r, r, part);
                    } else {  // EPI_CHEB
                        const double r = P.a0[idx] - au;
                        const double dn = P.s0 * P.o1[idx] + P.s1 * (double)P.a1f[idx] * r;
                        P.o1[idx] = dn;
                        P.y[idx] = uo[c] + dn;
                    }
                }
            }
            yoff += P.pstride;
            if (FLAGS) {  // this row is done with plane pidx and with the exchange slot of step t
                __syncwarp();
                if (tx == 0) mbar_arrive(&ebars[pidx % L::STAGES]);
            }
        };

        // ---- march -------------------------------------------------------------------------------
        const int nsteps = z1 - pfirst;
        if (!FLAGS || rowcompute) {
            for (int t = 0; t < nsteps; t++) {
                step(t, fA, fB, uA, uB, mA, mB);
#pragma unroll
                for (int k = 0; k < 12; k++) fA[k] = fB[k];
#pragma unroll
                for (int k = 0; k < 3; k++) uA[k] = uB[k];
                mA = mB;
            }
        }
    }
    if (EPI == EPI_DOT || EPI == EPI_RESID) {
        part = block_sum(part, red);
        if (tid == 0) P.partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = part;
    }
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------------------------------
// z-march, generation 2: the mbarrier hand-off kernel above with the march unrolled by STAGES (= 4).
// What the unrolling buys (SASS of the gen-1 loop: 366 instructions per warp-step for 177 DP -- issue slots, not the
// FP64 pipe, were the binding resource, profiles/r01_ncu_zmarch_ty16_mbar_v3_256.txt):
//   * the carried state (lower-face modes, own nodal value, mask) ping-pongs between two register sets by NAME, so the
//     ~33 IMAD.MOV per step of the gen-1 loop disappear;
//   * ring-slot index, exchange-slot index and the barrier addresses are compile-time constants of the sub-step: the
//     uniform-datapath index arithmetic (USHF/ULOP3/ULEA/UIADD3, ~28 per step) disappears, shared-memory operands
//     become base + immediate.
// Everything else (tile geometry, producer warp, row hand-off protocol) is the gen-1 design.
// ---------------------------------------------------------------------------------------
template <int K>
struct IntC { static constexpr int v = K; };

// MERGED = false: TY node rows per tile, warp row TY-1 is the dedicated TMA producer (15 compute warps of 16: one
//   scheduler runs 3 compute warps + the producer, the other three 4 compute warps -- those three bound the step).
// MERGED = true : TY + 1 node rows per tile, every warp owns an element row (4 compute warps on every scheduler, 15
//   instead of 14 output rows per step); the last row -- the tail of the hand-off wavefront, so the other rows have
//   released the slot by the time it gets there -- also refills the TMA ring.
template <int TY, int EPI, bool MERGED>
__global__ void __launch_bounds__(32 * TY, 1)
k_hex8_zm4(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmE,
           const __grid_constant__ Hex8BlocksPacked KB, const __grid_constant__ ZmParams P) {
    constexpr int ROWS = MERGED ? TY + 1 : TY;   // node rows in the box
    constexpr int NCOMP = ROWS - 1;              // element rows = compute warps
    using L = ZmSmem<TY, true, ROWS>;
    static_assert(L::STAGES == 4 && L::XCH_DEPTH == 4, "the unrolled march assumes 4 ring slots");
    constexpr int NT = L::NT;
    constexpr int NC = L::NC;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double red[32];
    double* xch = reinterpret_cast<double*>(smem_raw + L::STAGES * L::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::BAR_OFF);  // full[4]
    uint64_t* ebars = bars + L::STAGES;                                   // empty[4]
    uint64_t* xbars = bars + 2 * L::STAGES;                               // xfull[TY][4]

    const int tid = threadIdx.x;
    const int tx = tid & 31;
    const int ty = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int i0 = blockIdx.x * 30 - 1, j0 = blockIdx.y * (ROWS - 2) - 1;
    const int i = i0 + tx, j = j0 + ty;
    const int z0 = P.zout0 + blockIdx.z * P.zchunk;
    const int z1 = min(z0 + P.zchunk, P.zout1);
    const bool outthr = (tx >= 1) && (tx <= 30) && (ty >= 1) && (ty <= ROWS - 2) && (i < P.nxn) && (j < P.nyn);
    const int pfirst = z0 - 1;
    const int nplanes = z1 - z0 + 2;
    double part = 0.0;
    // Rows whose node row lies outside the grid (j = -1 in the first tile row, j >= nyn in the last) have nothing to
    // compute and nothing to write: they stay out of the march altogether (fewer warps on the FP64 pipe in the edge tiles).
    const int ja = max(j0, 0), jb = min(j0 + NCOMP - 1, P.nyn - 1);  // active element rows: node rows ja..jb
    const int nactive = jb - ja + 1;
    const bool rowactive = (j >= ja) && (j <= jb) && (ty < NCOMP);

    if (EPI == EPI_CGDOT) pdl_launch_dependents();  // the next U kernel may become resident as CTAs retire (it parks at its own wait)
    if (z0 < z1) {
        if (tid == 0) {
            // programmatic dependent launch: this grid may have started while the U kernel was draining -- thread 0 waits for its
            // completion here; every other thread's first global access lies behind the block barrier below
            if (EPI == EPI_CGDOT) pdl_wait();
            if (EPI == EPI_CGDOT && (P.pe.n > 1 || P.pe.trace)) {
                // the boundary planes of the operand are stored into my ghost planes by the neighbours' U kernels: only the
                // chunks that read a ghost plane wait (system-scope acquire of the epoch flag in my arena)
                const unsigned long long e = P.cg->st[P.cgpar].epoch;
                const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(P.pe.arena);
                const bool tr = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == (P.pe.rank > 0 ? 0 : gridDim.z - 1));
                if (tr) cg_trace(P.pe, e, 3);
                const long long t0 = clock64();  // bounded: a dead neighbour must not hang this GPU (the solve then fails its residual check)
                if (z0 == P.zout0 && P.pe.rank > 0)
                    while (ld_acquire_sys_u64(&flags[F_CG_FROM_LO]) < e && clock64() - t0 < 8000000000LL) {}
                if (z1 == P.zout1 && P.pe.rank < P.pe.n - 1)
                    while (ld_acquire_sys_u64(&flags[F_CG_FROM_HI]) < e && clock64() - t0 < 8000000000LL) {}
                if (tr) cg_trace(P.pe, e, 4);
            }
            for (int s = 0; s < L::STAGES; s++) mbar_init(&bars[s], 1);
            for (int s = 0; s < L::STAGES; s++) mbar_init(&ebars[s], nactive);
            for (int s = 0; s < TY * L::XCH_DEPTH; s++) mbar_init(&xbars[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();

        auto issue = [&](const int q) {  // plane counter q -> ring slot q & 3
            const int s = q & 3;
            unsigned char* st = smem_raw + s * L::STAGE_BYTES;
            const int p = pfirst + q;
            mbar_expect_tx(&bars[s], L::TX_BYTES);
            if (P.l2hint & 2) tma_load_4d_hint(st, &tmU, &bars[s], i0 - 1, j0, 0, p, l2_policy_last());
            else tma_load_4d(st, &tmU, &bars[s], i0 - 1, j0, 0, p);
            if (P.l2hint & 1) tma_load_3d_hint(st + L::U_BYTES, &tmE, &bars[s], i0 - 1, j0, p - 1, l2_policy_first());
            else tma_load_3d(st + L::U_BYTES, &tmE, &bars[s], i0 - 1, j0, p - 1);
        };
        const bool refills = MERGED && j == jb;  // the last ACTIVE row
        if (!MERGED && ty == TY - 1) {
            // ---- dedicated producer row ----------------------------------------------------------
            for (int q = 0; q < nplanes; q++) {
                if (q >= L::STAGES) mbar_wait(&ebars[q & 3], ((q >> 2) - 1) & 1);
                if (tx == 0) issue(q);
            }
        } else if (rowactive) {
            if (refills && tx == 0)
                for (int q = 0; q < L::STAGES && q < nplanes; q++) issue(q);
            // ---- consumer rows ---------------------------------------------------------------
            const bool innode = (i >= 0) && (i < P.nxn) && (j >= 0) && (j < P.nyn);
            const unsigned char* mptr = P.mask + ((long long)pfirst * P.mstride + (long long)j * P.pm + i);
            const int o = ty * ZM_W + tx + 1;
            const int o2 = ty * 32 + tx;
            double* ybase = P.y + ((long long)pfirst * P.pstride + (long long)j * P.px + i);  // EPI operands share the offset
            const long long adiff = (EPI == EPI_RESID || EPI == EPI_CHEB) ? (P.a0 - P.y) : 0;
            const double* su0 = reinterpret_cast<const double*>(smem_raw) + o;  // my node in slot 0, component 0
            double* xw0 = xch + o2;
            uint64_t* xb_mine = &xbars[ty * 4];
            uint64_t* xb_below = &xbars[(ty - 1) * 4];
            const bool has_below = j > ja, signals = j < jb;

            auto face_modes = [&](const double* su, double (&fm)[12], double (&own)[3]) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double a = su[c * NC], b = su[c * NC + 1], cc = su[c * NC + ZM_W], d = su[c * NC + ZM_W + 1];
                    const double s0 = a + b, d0 = a - b, s1 = cc + d, d1 = cc - d;
                    fm[4 * c + 0] = s0 + s1;
                    fm[4 * c + 1] = d0 + d1;
                    fm[4 * c + 2] = s0 - s1;
                    fm[4 * c + 3] = d0 - d1;
                    own[c] = a;
                }
            };

            double fA[12], fB[12], upr[12], uA[3], uB[3];
            unsigned mA = 0, mB = 0;
#pragma unroll
            for (int k = 0; k < 12; k++) { fA[k] = 0.0; fB[k] = 0.0; upr[k] = 0.0; }
#pragma unroll
            for (int c = 0; c < 3; c++) { uA[c] = 0.0; uB[c] = 0.0; }

            // prologue: plane pfirst (ring slot 0, phase 0)
            mbar_wait(&bars[0], 0);
            mA = (innode && pfirst >= 0) ? (unsigned)*mptr : 0u;
            mptr += P.mstride;
            face_modes(su0, fA, uA);
            __syncwarp();
            if (tx == 0) mbar_arrive(&ebars[0]);
            if (refills && L::STAGES < nplanes) {  // plane 4 reuses the slot of plane 0
                mbar_wait(&ebars[0], 0);
                if (tx == 0) issue(L::STAGES);
            }

            // sub-step K of a group of four: t = 4m + K, upper plane pidx = t + 1 sits in ring slot (K+1)&3
            auto step = [&](auto kc, const int t, const unsigned ph, const double (&fl)[12], double (&fn)[12], const double (&uo)[3],
                            double (&un)[3], const unsigned mo, unsigned& mn) {
                constexpr int K = decltype(kc)::v;
                constexpr int S = (K + 1) & 3;                       // ring slot of plane t + 1
                const unsigned php = (K == 3) ? (ph ^ 1u) : ph;      // its phase parity: ((t + 1) >> 2) & 1
                mn = (innode && pfirst + t + 1 < P.nzl) ? (unsigned)*mptr : 0u;
                mptr += P.mstride;
                mbar_wait(&bars[S], php);
                const double2* kz = &KB.k[0][0] + (t >> (30 - K));  // see gen-1: keeps the constant loads per-step LDCU; a different shift per sub-step so they are not merged
                const double* su = su0 + S * (L::STAGE_BYTES / 8);
                double* xw = xw0 + K * 6 * NT;

                double ex0[3][2];
                {
                    face_modes(su, fn, un);
                    const double Ee = su[L::U_BYTES / 8];
                    double uh[3][8];
#pragma unroll
                    for (int c = 0; c < 3; c++)
#pragma unroll
                        for (int f = 0; f < 4; f++) {
                            uh[c][f] = fl[4 * c + f] + fn[4 * c + f];
                            uh[c][4 + f] = fl[4 * c + f] - fn[4 * c + f];
                        }
                    double yh[3][8];
                    // Block rho couples (mode rho^(1<<c), component c), c = 0..2.  Entry (c,d), c != d, is non-zero only when
                    // bits c and d of rho are equal (both derivatives come from the lambda term or both from the mu term);
                    // (c,c) vanishes for rho == 1<<c (rigid translation).  So rho = 0,7 are full symmetric 3x3, rho = 1,2,4 a
                    // 2x2, rho = 3,5,6 a 2x2 plus a 1x1: 45 DP operations instead of 72 (checked in hex8_build_operator).
#pragma unroll
                    for (int rho = 0; rho < 8; rho++) {
                        const double in[3] = {uh[0][rho ^ 1], uh[1][rho ^ 2], uh[2][rho ^ 4]};
                        const double2 k0 = kz[rho * 3 + 0], k1 = kz[rho * 3 + 1], k2 = kz[rho * 3 + 2];
                        const double b[3][3] = {{k0.x, k0.y, k1.x}, {k0.y, k1.y, k2.x}, {k1.x, k2.x, k2.y}};
                        double o3[3];
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            double acc = 0.0;
                            bool first = true;
#pragma unroll
                            for (int d = 0; d < 3; d++) {
                                const bool active = (c == d) ? (rho != (1 << c)) : (((rho >> c) & 1) == ((rho >> d) & 1));
                                if (active) {
                                    acc = first ? b[c][d] * in[d] : fma(b[c][d], in[d], acc);
                                    first = false;
                                }
                            }
                            o3[c] = acc;
                        }
                        yh[0][rho ^ 1] = o3[0];
                        yh[1][rho ^ 2] = o3[1];
                        yh[2][rho ^ 4] = o3[2];
                    }
                    double ht[3][4];
#pragma unroll
                    for (int c = 0; c < 3; c++)
#pragma unroll
                        for (int f = 0; f < 4; f++) {
                            const double sl = yh[c][f] + yh[c][4 + f], su2 = yh[c][f] - yh[c][4 + f];
                            ht[c][f] = fma(Ee, sl, upr[4 * c + f]);
                            upr[4 * c + f] = Ee * su2;
                        }
#pragma unroll
                    for (int c = 0; c < 3; c++)
#pragma unroll
                        for (int ma = 0; ma < 2; ma++) {
                            ex0[c][ma] = ht[c][ma] + ht[c][ma + 2];
                            xw[(2 * c + ma) * NT] = ht[c][ma] - ht[c][ma + 2];
                        }
                }
                __syncwarp();
                if (tx == 0 && signals) mbar_arrive(&xb_mine[K]);
                if (has_below) mbar_wait(&xb_below[K], ph);

                double v[3], vr[3];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    double e0 = ex0[c][0], e1 = ex0[c][1];
                    if (has_below) {
                        e0 += xw[(2 * c + 0) * NT - 32];
                        e1 += xw[(2 * c + 1) * NT - 32];
                    }
                    v[c] = e0 + e1;
                    vr[c] = e0 - e1;
                }
#pragma unroll
                for (int c = 0; c < 3; c++) v[c] += __shfl_up_sync(0xffffffffu, vr[c], 1);

                const int e = pfirst + t;
                if (outthr && e >= z0) {
                    double* yp = ybase;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const bool fixed = (mo >> c) & 1u;
                        const double au = fixed ? uo[c] : v[c];
                        if (EPI == EPI_Y) {
                            yp[c * P.nstride] = au;
                        } else if (EPI == EPI_DOT || EPI == EPI_CGDOT) {
                            yp[c * P.nstride] = au;
                            part = fma(uo[c], au, part);
                        } else if (EPI == EPI_RESID) {
                            const double r = yp[c * P.nstride + adiff] - au;
                            yp[c * P.nstride] = r;
                            part = fma(r, r, part);
                        } else {  // EPI_CHEB
                            const long long idx = (yp - P.y) + c * P.nstride;
                            const double r = P.a0[idx] - au;
                            const double dn = P.s0 * P.o1[idx] + P.s1 * (double)P.a1f[idx] * r;
                            P.o1[idx] = dn;
                            yp[c * P.nstride] = uo[c] + dn;
                        }
                    }
                }
                ybase += P.pstride;
                __syncwarp();
                if (tx == 0) mbar_arrive(&ebars[S]);
                if (MERGED) {
                    if (refills && t + 1 + L::STAGES < nplanes) {  // every row is done with plane t + 1: refill its slot
                        mbar_wait(&ebars[S], php);
                        if (tx == 0) issue(t + 1 + L::STAGES);
                    }
                }
            };

            const int nsteps = z1 - pfirst;
            unsigned ph = 0;
            for (int t0 = 0; t0 < nsteps; t0 += 4, ph ^= 1u) {
                step(IntC<0>(), t0, ph, fA, fB, uA, uB, mA, mB);
                if (t0 + 1 >= nsteps) break;
                step(IntC<1>(), t0 + 1, ph, fB, fA, uB, uA, mB, mA);
                if (t0 + 2 >= nsteps) break;
                step(IntC<2>(), t0 + 2, ph, fA, fB, uA, uB, mA, mB);
                if (t0 + 3 >= nsteps) break;
                step(IntC<3>(), t0 + 3, ph, fB, fA, uB, uA, mB, mA);
            }
        }
    }
    if (EPI == EPI_DOT || EPI == EPI_RESID || EPI == EPI_CGDOT) {
        part = block_sum(part, red);
        if (tid == 0) P.partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = part;
    }
    if (EPI == EPI_CGDOT) {
        // the last CTA to finish reduces every partial of this iteration (U's r.u and r.r, this kernel's w.u) in a fixed
        // order and publishes the three sums: locally on one rank, into every peer's arena (+ epoch flag) otherwise
        __shared__ int s_last;
        const unsigned int nblk = gridDim.x * gridDim.y * gridDim.z;
        if (tid == 0) {
            __threadfence();
            s_last = (atomicAdd(&P.cg->scounter, 1u) == nblk - 1u);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double a = 0.0, b = 0.0, d = 0.0;
            for (int k = tid; k < P.nbU; k += blockDim.x) { a += __ldcg(P.upart + k); b += __ldcg(P.upart + P.nbU + k); }
            for (unsigned int k = tid; k < nblk; k += blockDim.x) d += __ldcg(P.partials + k);
            a = block_sum(a, red);
            b = block_sum(b, red);
            d = block_sum(d, red);
            __shared__ double s_sum[3];
            if (tid == 0) { s_sum[0] = a; s_sum[1] = b; s_sum[2] = d; P.cg->scounter = 0u; }
            __syncthreads();
            if (P.pe.n == 1) {
                if (tid < 3) P.cg->sums[tid] = s_sum[tid];
                if (tid == 0) cg_trace(P.pe, P.cg->st[P.cgpar].epoch, 5);
            } else {
                const unsigned long long e = P.cg->st[P.cgpar].epoch;
                const size_t off = ((e & 1) ? P.red_off1 : P.red_off0) + (size_t)P.pe.rank * 128 * 8;
                if (tid < P.pe.n) {
                    double* dst = reinterpret_cast<double*>(P.pe.peer[tid] + off) + kCgSlot;
                    dst[0] = s_sum[0]; dst[1] = s_sum[1]; dst[2] = s_sum[2];
                    fence_acq_rel_sys();  // sums -> fence -> relaxed flag store (one system fence instead of two)
                    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(P.pe.peer[tid]) + F_CG_RED0 + P.pe.rank, e);
                }
                if (tid == 0) cg_trace(P.pe, e, 5);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// host launch
// ---------------------------------------------------------------------------------------
template <int TY, int EPI, bool FLAGS, int GEN = 1>
static int launch_zm(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const SpmvArgs& a, int* nparts) {
    constexpr int ROWS = (GEN == 3) ? TY + 1 : TY;  // node rows of the TMA box (gen 3 = gen 2 with the producer merged into the last row)
    using L = ZmSmem<TY, FLAGS, ROWS>;
    CUtensorMap tmU, tmE;
    {
        cuuint64_t dims[4] = {(cuuint64_t)g.nxn, (cuuint64_t)g.nyn, 3, (cuuint64_t)g.nzl()};
        cuuint64_t str[3] = {(cuuint64_t)g.px * 8, (cuuint64_t)g.nstride * 8, (cuuint64_t)g.pstride * 8};
        cuuint32_t box[4] = {ZM_W, (cuuint32_t)ROWS, 3, 1};
        TOB_CHECK(encode_map(c, &tmU, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, a.u, dims, str, box));
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)g.nex, (cuuint64_t)g.ney, (cuuint64_t)g.nzle()};
        cuuint64_t str[2] = {(cuuint64_t)g.pe * 8, (cuuint64_t)g.estride * 8};
        cuuint32_t box[3] = {ZM_W, (cuuint32_t)ROWS, 1};
        TOB_CHECK(encode_map(c, &tmE, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, a.E, dims, str, box));
    }
    ZmParams P;
    P.nxn = g.nxn; P.nyn = g.nyn; P.nzl = g.nzl(); P.px = g.px; P.pm = g.pm; P.mstride = g.mstride; P.mask = a.mask; P.nstride = g.nstride; P.pstride = g.pstride;
    P.zout0 = g.zl_own0(); P.zout1 = g.zl_own1();
    P.y = a.y; P.a0 = a.a0; P.a1 = a.a1; P.a1f = a.a1f; P.o1 = a.o1; P.s0 = a.s0; P.s1 = a.s1; P.partials = c->d_partials;
    P.l2hint = a.l2hint;
    P.cg = a.cg; P.cgpar = a.cgpar; P.upart = a.upart; P.nbU = a.nbU;
    if (a.pe) P.pe = *a.pe; else P.pe = CgPeers();
    P.red_off0 = P.pe.n > 1 ? cg_red_off(P.pe.n, 0, 0) : 0;
    P.red_off1 = P.pe.n > 1 ? cg_red_off(P.pe.n, 1, 0) : 0;
    const int gx = div_up(g.nxn, 30), gy = div_up(g.nyn, ROWS - 2);
    const int nz = P.zout1 - P.zout0;
    // z-chunks: every chunk recomputes one element layer and pays a prologue, and the grid runs in waves
    // of ctas_per_wave CTAs (ncu: with 3.16 waves the SMs idled 21% of the time) -> pick the chunk count
    // that minimises  waves * (planes per chunk + overhead).
    const int ctas_per_wave = c->num_sms * ((TY <= 8) ? 2 : 1);
    int nch = 1;
    {
        double best = 1e300;
        const long long tiles = (long long)gx * gy;
        for (int k = 1; k <= nz; k++) {
            const int zc = div_up(nz, k);
            if (zc < 4 && k > 1) break;
            const int kk = div_up(nz, zc);
            const double waves = ceil((double)(tiles * kk) / ctas_per_wave);
            const double cost = waves * (zc + 3.0);
            if (cost < best * 0.999) { best = cost; nch = kk; }
        }
    }
    P.zchunk = div_up(nz, nch);
    nch = div_up(nz, P.zchunk);
    if ((long long)gx * gy * nch > kMaxPartials) TOB_FAIL("grid too large for the partial-sum buffer");
    void (*kern)(const CUtensorMap, const CUtensorMap, const Hex8BlocksPacked, const ZmParams);
    if constexpr (GEN == 2) kern = k_hex8_zm4<TY, EPI, false>;
    else if constexpr (GEN == 3) kern = k_hex8_zm4<TY, EPI, true>;
    else kern = k_hex8_zmarch<TY, EPI, FLAGS>;
    // the opt-in is per device (a process may hold contexts on several devices): cheap, so set it on every launch path once per device
    static bool attr_set[64] = {};
    if (c->device < 0 || c->device >= 64 || !attr_set[c->device]) {
        TOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        if (c->device >= 0 && c->device < 64) attr_set[c->device] = true;
    }
    Hex8BlocksPacked kp;
    for (int rho = 0; rho < 8; rho++) {
        const double* b = op.blocks.b[rho];
        kp.k[rho][0] = make_double2(b[0], b[1]);
        kp.k[rho][1] = make_double2(b[2], b[4]);
        kp.k[rho][2] = make_double2(b[5], b[8]);
    }
    if (EPI == EPI_CGDOT && pdl_enabled()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gx, gy, nch); cfg.blockDim = dim3(32 * TY); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        TOB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmU, tmE, kp, P));
    } else {
        kern<<<dim3(gx, gy, nch), 32 * TY, L::TOTAL, c->stream>>>(tmU, tmE, kp, P);
    }
    c->launches++;
    TOB_KERNEL_CHECK();
    if (nparts) *nparts = (EPI == EPI_DOT || EPI == EPI_RESID || EPI == EPI_CGDOT) ? gx * gy * nch : 0;
    return 0;
}

template <int TY, bool FLAGS, int GEN = 1>
static int dispatch_zm(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const SpmvArgs& a, int* nparts) {
    switch (a.epi) {
        case EPI_Y: return launch_zm<TY, EPI_Y, FLAGS, GEN>(c, g, op, a, nparts);
        case EPI_DOT: return launch_zm<TY, EPI_DOT, FLAGS, GEN>(c, g, op, a, nparts);
        case EPI_RESID: return launch_zm<TY, EPI_RESID, FLAGS, GEN>(c, g, op, a, nparts);
        case EPI_CGDOT:
            if constexpr (GEN == 2 && TY == 16) return launch_zm<TY, EPI_CGDOT, FLAGS, GEN>(c, g, op, a, nparts);
            else TOB_FAIL("the single-reduction PCG epilogue exists for the default SpMV kernel only");
        default: return launch_zm<TY, EPI_CHEB, FLAGS, GEN>(c, g, op, a, nparts);
    }
}

int hex8_spmv(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const SpmvArgs& a, int* nparts) {
    if (nparts) *nparts = 0;
    // measured (profiles/r01_spmv_probe_v6.txt): the mbarrier hand-off kernel with 32x16 tiles wins at every size tried
    // (192x96x96: 43 us vs 49 (same tile, __syncthreads) / 57-61 (32x8 tiles); 256^3: 0.325 ms vs 0.385 / 0.362-0.391)
    int kernel = a.kernel == SPMV_AUTO ? SPMV_ZM16U4 : a.kernel;  // gen 2 (profiles/r02_spmv_probe_*.txt)
    if (a.mask_in && kernel != SPMV_NAIVE) TOB_FAIL("z-march kernels expect a pre-masked input (N.u); mask the input first");
    if (a.epi == EPI_CHEB && !a.a1f) TOB_FAIL("Chebyshev epilogue: the fp32 copy of the inverse diagonal (SpmvArgs::a1f) is missing");
    if (kernel == SPMV_NAIVE) {
        if (a.epi != EPI_Y) TOB_FAIL("naive SpMV kernel supports only the plain epilogue");
        dim3 grid(div_up(g.nxn, 128), g.nyn, g.zl_own1() - g.zl_own0());
        k_hex8_naive<<<grid, 128, 0, c->stream>>>(a.u, a.y, a.E, a.mask, op.d_KE, g.nxn, g.nyn, g.nzl(), g.px, g.pe, g.pm,
                                                  g.nstride, g.pstride, g.estride, g.mstride, g.zl_own0(), g.zl_own1(),
                                                  a.mask_in ? 1 : 0);
        c->launches++;
        TOB_KERNEL_CHECK();
        return 0;
    }
    // (the shared-memory-state variants SPMV_ZM8 / SPMV_ZM16 of the first kernel generation are retired: their ids
    // run the register-state kernels of the same tile shape)
    if (kernel == SPMV_ZM16 || kernel == SPMV_ZM16R) return dispatch_zm<16, false>(c, g, op, a, nparts);
    if (kernel == SPMV_ZM12) return dispatch_zm<12, false>(c, g, op, a, nparts);
    if (kernel == SPMV_ZM8RF) return dispatch_zm<8, true>(c, g, op, a, nparts);
    if (kernel == SPMV_ZM16U4) return dispatch_zm<16, true, 2>(c, g, op, a, nparts);
    if (kernel == SPMV_ZM16M4) return dispatch_zm<16, true, 3>(c, g, op, a, nparts);
    if (kernel == SPMV_ZM16RF) return dispatch_zm<16, true>(c, g, op, a, nparts);
    if (kernel == SPMV_ZM12RF) return dispatch_zm<12, true>(c, g, op, a, nparts);
    return dispatch_zm<8, false>(c, g, op, a, nparts);  // SPMV_ZM8, SPMV_ZM8R
}

// ---------------------------------------------------------------------------------------
// diagonal, SIMP modulus, compliance / sensitivities
// ---------------------------------------------------------------------------------------
__global__ void k_hex8_diag(const double* __restrict__ E, const unsigned char* __restrict__ mask, double* __restrict__ d,
                            double kd0, double kd1, double kd2, int nxn, int nyn, int nzl, int px, int pe, int pm,
                            long long nstride, long long pstride, long long estride, long long mstride, int z0, int z1,
                            int invert) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int z = z0 + blockIdx.z;
    if (i >= nxn || z >= z1) return;
    const int nex = nxn - 1, ney = nyn - 1, nzle = nzl - 1;
    double s = 0.0;
    for (int dk = -1; dk <= 0; dk++)
        for (int dj = -1; dj <= 0; dj++)
            for (int di = -1; di <= 0; di++) {
                const int ei = i + di, ej = j + dj, ek = z + dk;
                if (ei < 0 || ei >= nex || ej < 0 || ej >= ney || ek < 0 || ek >= nzle) continue;
                s += E[ek * estride + (long long)ej * pe + ei];
            }
    const unsigned char m = mask[z * mstride + (long long)j * pm + i];
    const double kd[3] = {kd0, kd1, kd2};
    for (int c = 0; c < 3; c++) {
        double v = (m & (1 << c)) ? 1.0 : kd[c] * s;
        d[z * pstride + c * nstride + (long long)j * px + i] = invert ? 1.0 / v : v;
    }
}

int hex8_diag_inv(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const double* E,
                  const unsigned char* mask, double* d, bool invert) {
    dim3 grid(div_up(g.nxn, 128), g.nyn, g.zl_own1() - g.zl_own0());
    k_hex8_diag<<<grid, 128, 0, c->stream>>>(E, mask, d, op.kdiag[0], op.kdiag[1], op.kdiag[2], g.nxn, g.nyn, g.nzl(), g.px,
                                             g.pe, g.pm, g.nstride, g.pstride, g.estride, g.mstride, g.zl_own0(),
                                             g.zl_own1(), invert ? 1 : 0);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

__global__ void __launch_bounds__(kFlatThreads) k_simp(const double* __restrict__ x, double* __restrict__ E, long long n,
                                                       double Emin, double Emax, double penal) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        E[i] = Emin + pow(x[i], penal) * (Emax - Emin);
}
int simp_modulus(tob_ctx_impl* c, const double* x, double* E, long long n, double Emin, double Emax, double penal) {
    k_simp<<<flat_grid(c, n), kFlatThreads, 0, c->stream>>>(x, E, n, Emin, Emax, penal);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

// one thread per owned element: uKu = uh^T D uh in the symmetry basis
__global__ void __launch_bounds__(128)
k_hex8_compliance(const double* __restrict__ u, const double* __restrict__ xphys, const double* __restrict__ E,
                  double* __restrict__ dfdx, const __grid_constant__ Hex8Blocks KB, int nex, int ney, int px, int pe,
                  long long nstride, long long pstride, long long estride, int l0, int l1, double penal, double dE,
                  double* __restrict__ partials) {
    __shared__ double red[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    double pf = 0.0, px_sum = 0.0;
    // grid.z may be smaller than the number of layers (partial-sum buffer): every CTA strides over layers
    for (int l = l0 + blockIdx.z; i < nex && l < l1; l += gridDim.z) {
        double uh[3][8];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double w[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
                w[q] = u[(l + (q >> 2)) * pstride + c * nstride + (long long)(j + ((q >> 1) & 1)) * px + i + (q & 1)];
#pragma unroll
            for (int h = 1; h < 8; h <<= 1)
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (!(q & h)) { const double a = w[q], b = w[q | h]; w[q] = a + b; w[q | h] = a - b; }
#pragma unroll
            for (int q = 0; q < 8; q++) uh[c][q] = w[q];
        }
        double uKu = 0.0;
#pragma unroll
        for (int rho = 0; rho < 8; rho++) {
            const double in[3] = {uh[0][rho ^ 1], uh[1][rho ^ 2], uh[2][rho ^ 4]};
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) uKu = fma(in[c] * KB.b[rho][3 * c + c2], in[c2], uKu);
        }
        const long long eidx = l * estride + (long long)j * pe + i;
        const double xp = xphys[eidx];
        pf += E[eidx] * uKu;
        px_sum += xp;
        dfdx[(long long)i + (long long)nex * (j + (long long)ney * (l - l0))] = -penal * pow(xp, penal - 1.0) * dE * uKu;
    }
    pf = block_sum(pf, red);
    px_sum = block_sum(px_sum, red);
    if (threadIdx.x == 0) {
        const long long nb = (long long)gridDim.x * gridDim.y * gridDim.z;
        const long long b = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partials[b] = pf;
        partials[nb + b] = px_sum;
    }
}

int hex8_compliance(tob_ctx_impl* c, const LevelGrid& g, const Hex8Operator& op, const double* u, const double* xphys,
                    const double* E, double penal, double Emin, double Emax, double* dfdx_nat, int slot) {
    const int l0 = g.ke0 - g.ge0, l1 = g.ke1 - g.ge0;
    dim3 grid(div_up(g.nex, 128), g.ney, l1 - l0);
    // two partial sums per CTA must fit the buffer: on large slabs the CTAs stride over the layers instead
    const long long per_layer = (long long)grid.x * grid.y;
    if (2 * per_layer > kMaxPartials) TOB_FAIL("compliance: one element layer exceeds the partial-sum buffer");
    grid.z = (unsigned)std::max(1LL, std::min<long long>(l1 - l0, kMaxPartials / 2 / per_layer));
    const long long nb = (long long)grid.x * grid.y * grid.z;
    k_hex8_compliance<<<grid, 128, 0, c->stream>>>(u, xphys, E, dfdx_nat, op.blocks, g.nex, g.ney, g.px, g.pe, g.nstride,
                                                   g.pstride, g.estride, l0, l1, penal, Emax - Emin, c->d_partials);
    c->launches++;
    TOB_KERNEL_CHECK();
    return finalize_sum(c, (int)nb, 2, slot, true);
}

}  // namespace tob
