// filter.cu -- density / sensitivity cone filter as a stencil, Heaviside projection, MND.
// Replaces class Filter (Filter.cc:25-519): the explicit AIJ matrix H (Filter.cc:404-440) is never
// formed -- its weights depend only on the index offset, so H x is a (2s+1)^3 convolution with a
// clipped box, and Hs = H 1 (Filter.cc:445-448) is the same sum with x = 1.
#include "filter.cuh"
#include "tma.cuh"
#include "vecops.cuh"

namespace tob {

enum { F_S_MND = 40, F_S_TMP = 42 };

// ---- layout helpers: natural owned element vector <-> filter field with s ghost layers -------
// mode 0: st = a ; 1: st = a / hs_st ; 2: st = a * b
__global__ void k_nat_to_filter(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ hs,
                                double* __restrict__ st, int nex, int ney, int nzo, int pe, long long estride, int own0,
                                int mode) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)nex * ney * nzo;
    if (idx >= tot) return;
    int i = idx % nex; long long t = idx / nex;
    int j = t % ney; int k = t / ney;
    const long long s = (k + own0) * estride + (long long)j * pe + i;
    double v = a[idx];
    if (mode == 1) v /= hs[s];
    else if (mode == 2) v *= b[idx];
    st[s] = v;
}

// out(natural) = stencil sum of `in` (filter layout), mode 0: / sum w ; 1: plain ; 2: / sum w / x
// mode 3: out(filter layout) = sum w  (Hs)
__global__ void __launch_bounds__(128)
k_cone(const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ wtab,
       const double* __restrict__ xnat, int nex, int ney, int nez, int pe, long long estride, int s, int gl0, int l0,
       int l1, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int l = l0 + blockIdx.z;  // local layer
    if (i >= nex || l >= l1) return;
    const int kg = gl0 + l;  // global layer
    const int w1 = 2 * s + 1;
    double acc = 0.0, wsum = 0.0;
    const int dk0 = max(-s, -kg), dk1 = min(s, nez - 1 - kg);
    const int dj0 = max(-s, -j), dj1 = min(s, ney - 1 - j);
    const int di0 = max(-s, -i), di1 = min(s, nex - 1 - i);
    for (int dk = dk0; dk <= dk1; dk++)
        for (int dj = dj0; dj <= dj1; dj++) {
            const double* row = in + (l + dk) * estride + (long long)(j + dj) * pe + i;
            const double* wr = wtab + ((dk + s) * w1 + (dj + s)) * w1 + s;
            for (int di = di0; di <= di1; di++) {
                const double w = __ldg(wr + di);
                wsum += w;
                if (mode != 3) acc = fma(w, row[di], acc);
            }
        }
    if (mode == 3) {
        out[l * estride + (long long)j * pe + i] = wsum;
        return;
    }
    const long long nidx = (long long)i + (long long)nex * (j + (long long)ney * (l - l0));
    double v = acc;
    if (mode == 0) v = acc / wsum;
    else if (mode == 2) v = acc / wsum / xnat[nidx];
    out[nidx] = v;
}

// ---- TMA-tiled cone filter (modes 0-2, s = 1, 2 or 3) ------------------------------------------------------------
// One CTA produces a 32 x 8 x CT_Z block of outputs.  A single TMA box load brings the block plus its s-wide rim
// into shared memory; everything outside the stored field (domain boundary, pitch padding, layers beyond the ghost
// range) is zero-filled by the TMA unit, which is exactly the clipped box of Filter.cc:417-419 for the weighted sum
// (the matching weight sum Hs is precomputed once, mode 3 of k_cone).  Each thread owns one (i,j) column: it walks
// the CT_Z + 2s planes of the box once, loads each of the (2s+1)^2 neighbours of a plane ONCE and scatters it into
// the up to 2s+1 output layers it belongs to (5x fewer shared-memory reads than a gather per output); the (s+1)^3
// distinct weights (they depend on |di|,|dj|,|dk| only) live in registers.
// HBM traffic: the field once (+ rim), Hs once, output once  ~ 24-32 B/element instead of (2s+1)^3 cached reads.
constexpr int CT_X = 32, CT_Y = 8, CT_Z = 8;
struct ConeTileParams {
    const double* hs;    // Hs, filter layout
    const double* xnat;  // mode 2: x, natural order
    const double* wtab;  // (2s+1)^3 weights
    double* out;         // natural order (owned layers)
    int nex, ney, pe, l0, l1, mode;
    long long estride;
};
template <int S>
struct ConeTile {
    static constexpr int SP = (S + 1) & ~1;        // x rim padded to an even count: 16-byte aligned TMA start column
    static constexpr int BW = CT_X + 2 * SP, BH = CT_Y + 2 * S, BD = CT_Z + 2 * S;
    static constexpr int BOX_BYTES = BW * BH * BD * 8;
    static constexpr int SMEM = BOX_BYTES + 128;   // + mbarrier
};

template <int S>
__global__ void __launch_bounds__(CT_X * CT_Y)
k_cone_tiled(const __grid_constant__ CUtensorMap tm, const __grid_constant__ ConeTileParams P) {
    using T = ConeTile<S>;
    extern __shared__ __align__(128) unsigned char cone_smem[];
    const double* box = reinterpret_cast<const double*>(cone_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(cone_smem + T::BOX_BYTES);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i0 = blockIdx.x * CT_X, j0 = blockIdx.y * CT_Y, l0 = P.l0 + blockIdx.z * CT_Z;
    if (tx == 0 && ty == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, T::BOX_BYTES);
        tma_load_3d(cone_smem, &tm, bar, i0 - T::SP, j0 - S, l0 - S);
    }
    // distinct weights w[|dk|][|dj|][|di|] while the box is in flight
    constexpr int W1 = 2 * S + 1;
    double w[S + 1][S + 1][S + 1];
#pragma unroll
    for (int a = 0; a <= S; a++)
#pragma unroll
        for (int b = 0; b <= S; b++)
#pragma unroll
            for (int c = 0; c <= S; c++) w[a][b][c] = __ldg(P.wtab + ((a + S) * W1 + (b + S)) * W1 + (c + S));
    __syncthreads();  // the barrier is initialised before anybody polls it
    mbar_wait(bar, 0);

    double acc[CT_Z];
#pragma unroll
    for (int o = 0; o < CT_Z; o++) acc[o] = 0.0;
    const double* col = box + (ty + S) * T::BW + (tx + T::SP);
#pragma unroll
    for (int zp = 0; zp < T::BD; zp++) {  // box plane zp holds local layer l0 - S + zp
#pragma unroll
        for (int dj = -S; dj <= S; dj++)
#pragma unroll
            for (int di = -S; di <= S; di++) {
                const double v = col[(zp * T::BH + dj) * T::BW + di];
#pragma unroll
                for (int dk = -S; dk <= S; dk++) {
                    const int o = zp - S - dk;  // output layer (tile-local) that sees this plane at offset dk
                    if (o >= 0 && o < CT_Z) acc[o] = fma(w[dk < 0 ? -dk : dk][dj < 0 ? -dj : dj][di < 0 ? -di : di], v, acc[o]);
                }
            }
    }
    const int i = i0 + tx, j = j0 + ty;
    if (i >= P.nex || j >= P.ney) return;
#pragma unroll
    for (int o = 0; o < CT_Z; o++) {
        const int l = l0 + o;
        if (l >= P.l1) break;
        const long long nidx = (long long)i + (long long)P.nex * (j + (long long)P.ney * (l - P.l0));
        double v = acc[o];
        if (P.mode != 1) v /= P.hs[l * P.estride + (long long)j * P.pe + i];
        if (P.mode == 2) v /= P.xnat[nidx];
        P.out[nidx] = v;
    }
}

template <int S>
static int cone_apply_tiled(tob_filter_impl* h, const double* in_st, double* out, const double* xnat, int mode) {
    using T = ConeTile<S>;
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = c->g0;
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)g.nex, (cuuint64_t)g.ney, (cuuint64_t)h->nl};
    cuuint64_t str[2] = {(cuuint64_t)g.pe * 8, (cuuint64_t)h->estride * 8};
    cuuint32_t bx[3] = {(cuuint32_t)T::BW, (cuuint32_t)T::BH, (cuuint32_t)T::BD};
    TOB_CHECK(encode_map(c, &tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, in_st, dims, str, bx));
    ConeTileParams P;
    P.hs = h->hs_st; P.xnat = xnat; P.wtab = h->wtab; P.out = out;
    P.nex = g.nex; P.ney = g.ney; P.pe = g.pe; P.l0 = h->own0; P.l1 = h->own1; P.mode = mode; P.estride = h->estride;
    auto kern = k_cone_tiled<S>;
    static bool attr_set[64] = {};  // the opt-in is per device
    if (c->device < 0 || c->device >= 64 || !attr_set[c->device]) {
        TOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM));
        if (c->device >= 0 && c->device < 64) attr_set[c->device] = true;
    }
    dim3 grid(div_up(g.nex, CT_X), div_up(g.ney, CT_Y), div_up(h->own1 - h->own0, CT_Z));
    kern<<<grid, dim3(CT_X, CT_Y), T::SMEM, c->stream>>>(tm, P);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

static int fhalo(tob_filter_impl* h, double* v) {
    tob_ctx_impl* c = h->c;
    if (c->desc.nranks == 1) return 0;
    const LevelGrid& g = c->g0;
    const int s = h->s;
    // what the neighbours store of my layers: lower neighbour's upper ghost width, upper neighbour's lower ghost width
    const int send_down = c->desc.rank > 0 ? min(s, g.nez - g.ke0) : 0;
    const int send_up = c->desc.rank < c->desc.nranks - 1 ? min(s, g.ke1) : 0;
    return halo_elems(c, v, h->estride, h->own0, h->own1, h->own0, h->nl - h->own1, send_down, send_up);
}

static int cone_apply(tob_filter_impl* h, const double* in_st, double* out, const double* xnat, int mode) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = c->g0;
    int l0 = h->own0, l1 = h->own1;
    if (mode != 3 && h->tiled) {
        if (h->s == 1) return cone_apply_tiled<1>(h, in_st, out, xnat, mode);
        if (h->s == 2) return cone_apply_tiled<2>(h, in_st, out, xnat, mode);
        if (h->s == 3) return cone_apply_tiled<3>(h, in_st, out, xnat, mode);  // 7^3 box: 64 distinct weights in registers, 1 CTA per SM
    }
    if (mode == 3) { l0 = 0; l1 = h->nl; }
    dim3 grid(div_up(g.nex, 128), g.ney, l1 - l0);
    k_cone<<<grid, 128, 0, c->stream>>>(in_st, out, h->wtab, xnat, g.nex, g.ney, g.nez, g.pe, h->estride, h->s, h->gl0, l0,
                                        l1, mode);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

static int to_filter(tob_filter_impl* h, const double* a, const double* b, double* st, int mode) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = c->g0;
    const int nzo = g.ke1 - g.ke0;
    long long tot = (long long)g.nex * g.ney * nzo;
    k_nat_to_filter<<<div_up(tot, 256), 256, 0, c->stream>>>(a, b, h->hs_st, st, g.nex, g.ney, nzo, g.pe, h->estride, h->own0, mode);
    c->launches++;
    TOB_KERNEL_CHECK();
    return fhalo(h, st);
}

// ---- pointwise: Heaviside projection (Filter.h:80-88), chain rule, MND ------------------------
__global__ void __launch_bounds__(kFlatThreads) k_heaviside(const double* __restrict__ x, double* __restrict__ y, long long n,
                                                            double beta, double eta) {
    const double den = tanh(beta * eta) + tanh(beta * (1.0 - eta));
    const double t0 = tanh(beta * eta);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = (t0 + tanh(beta * (x[i] - eta))) / den;
}
// g *= dH/dxTilde for up to 1+m gradient vectors
struct GradPtrs { double* p[9]; int n; };
__global__ void __launch_bounds__(kFlatThreads) k_heaviside_chain(const double* __restrict__ xt, GradPtrs gp, long long n,
                                                                  double beta, double eta) {
    const double den = tanh(beta * eta) + tanh(beta * (1.0 - eta));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double t = tanh(beta * (xt[i] - eta));
        const double d = beta * (1.0 - t * t) / den;
        for (int k = 0; k < gp.n; k++) gp.p[k][i] *= d;
    }
}
__global__ void __launch_bounds__(kFlatThreads) k_mnd(const double* __restrict__ x, long long n, double* __restrict__ partials) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += 4.0 * x[i] * (1.0 - x[i]);
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void __launch_bounds__(kFlatThreads) k_clamp01(double* __restrict__ x, long long n, double* __restrict__ partials) {
    // PDE filter bound check, Filter.cc:76-101: clamp to [0,1], report the largest violation
    __shared__ double red[32];
    double worst = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v < 0.0) { worst = fmax(worst, -v); v = 0.0; }
        if (v > 1.0) { worst = fmax(worst, v - 1.0); v = 1.0; }
        x[i] = v;
    }
    worst = block_max(worst, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = worst;
}

int filter_create(tob_ctx_impl* c, int type, double rmin, tob_filter_impl** out) {
    auto* h = new tob_filter_impl();
    h->c = c;
    h->type = type;
    h->R = rmin;
    const LevelGrid& g = c->g0;
    const long long ne = g.owned_elems();
    TOB_CUDA(cudaSetDevice(c->device));
    TOB_CUDA(cudaMalloc(&h->tmp_nat, sizeof(double) * ne));
    TOB_CUDA(cudaMalloc(&h->tmp_nat2, sizeof(double) * ne));
    if (type == 0 || type == 1) {
        // ElemConn, Filter.cc:326-327 (M,N,P are NODE counts)
        int s = (int)fmax(ceil(rmin / g.dx) - 1, fmax(ceil(rmin / g.dy) - 1, ceil(rmin / g.dz) - 1));
        s = std::min(s, std::min((g.nxn - 1) / 2, std::min((g.nyn - 1) / 2, (g.nzn - 1) / 2)));
        if (s < 0) s = 0;
        h->s = s;
        // The ghost layers come from the adjacent rank only, so every rank's slab must be at least s layers thick.  The
        // partition is known analytically on every rank: all ranks take the same decision, none is left waiting in a recv.
        for (int r = 0; r < c->desc.nranks && c->desc.nranks > 1; r++) {
            int e0 = 0, e1 = 0;
            TOB_CHECK(tob_slab_partition(g.nzn, c->desc.nranks, r, nullptr, nullptr, &e0, &e1));
            if (e1 - e0 < s) {
                delete h;
                TOB_FAIL("cone filter: the ghost width (ElemConn = " + std::to_string(s) + ") exceeds the slab thickness of rank " +
                         std::to_string(r) + " (" + std::to_string(e1 - e0) + " element layers); use fewer ranks or a smaller rmin");
            }
        }
        h->tiled = !(getenv("TOB_CONE_TILED") && atoi(getenv("TOB_CONE_TILED")) == 0);  // 0: direct-gather kernel (A/B, cross-check)
        const int lo = std::min(s, g.ke0), hi = std::min(s, g.nez - g.ke1);
        h->gl0 = g.ke0 - lo;
        h->own0 = lo;
        h->own1 = lo + (g.ke1 - g.ke0);
        h->nl = h->own1 + hi;
        h->estride = g.estride;
        const int w1 = 2 * s + 1;
        std::vector<double> w((size_t)w1 * w1 * w1);
        // weights R - dist for dist < R (Filter.cc:423-436); distances between uniform element centres
        for (int dk = -s; dk <= s; dk++)
            for (int dj = -s; dj <= s; dj++)
                for (int di = -s; di <= s; di++) {
                    const double ddx = di * g.dx, ddy = dj * g.dy, ddz = dk * g.dz;
                    const double dist = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
                    w[((size_t)(dk + s) * w1 + (dj + s)) * w1 + (di + s)] = dist < rmin ? rmin - dist : 0.0;
                }
        TOB_CUDA(cudaMalloc(&h->wtab, sizeof(double) * w.size()));
        TOB_CUDA(cudaMemcpyAsync(h->wtab, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, c->stream));
        const long long nst = h->estride * h->nl;
        TOB_CUDA(cudaMalloc(&h->f_st, sizeof(double) * nst));
        TOB_CUDA(cudaMalloc(&h->hs_st, sizeof(double) * nst));
        TOB_CUDA(cudaMemsetAsync(h->f_st, 0, sizeof(double) * nst, c->stream));
        TOB_CUDA(cudaMemsetAsync(h->hs_st, 0, sizeof(double) * nst, c->stream));
        TOB_CHECK(cone_apply(h, nullptr, h->hs_st, nullptr, 3));
        TOB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->desc.rank == 0 && 0) printf("# Filter radius rmin = %f results in a stencil of %i elements \n", rmin, s);
    } else if (type == 2) {
        TOB_CHECK(pdefilter_create(h));
    }
    *out = h;
    return 0;
}

void filter_destroy(tob_filter_impl* h) {
    if (!h) return;
    cudaSetDevice(h->c->device);
    cudaStreamSynchronize(h->c->stream);
    pdefilter_destroy(h);
    cudaFree(h->wtab); cudaFree(h->f_st); cudaFree(h->hs_st); cudaFree(h->tmp_nat); cudaFree(h->tmp_nat2);
    delete h;
}

}  // namespace tob

using namespace tob;

extern "C" {

int tob_filter_create(tob_ctx* ctx, int filter_type, double rmin, tob_filter** out) {
    if (!ctx || !out) TOB_FAIL("null argument");
    tob_filter_impl* h = nullptr;
    TOB_CHECK(filter_create(reinterpret_cast<tob_ctx_impl*>(ctx), filter_type, rmin, &h));
    *out = reinterpret_cast<tob_filter*>(h);
    return 0;
}
int tob_filter_destroy(tob_filter* h) {
    filter_destroy(reinterpret_cast<tob_filter_impl*>(h));
    return 0;
}
int tob_filter_elemconn(tob_filter* hh, int* s) {
    *s = reinterpret_cast<tob_filter_impl*>(hh)->s;
    return 0;
}

int tob_filter_project(tob_filter* hh, const double* x, double* xTilde, double* xPhys, int projection, double beta,
                       double eta) {
    auto* h = reinterpret_cast<tob_filter_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const long long ne = c->g0.owned_elems();
    const size_t nb = sizeof(double) * ne;
    const void* dx; void* dxt; void* dxp;
    TOB_CHECK(stage_in(c, 0, x, nb, &dx));
    TOB_CHECK(stage_out_begin(c, 1, xTilde, nb, &dxt));
    TOB_CHECK(stage_out_begin(c, 2, xPhys, nb, &dxp));
    if (h->type == 1) {  // xTilde = (H x)/Hs, Filter.cc:66-71
        TOB_CHECK(to_filter(h, (const double*)dx, nullptr, h->f_st, 0));
        TOB_CHECK(cone_apply(h, h->f_st, (double*)dxt, nullptr, 0));
    } else if (h->type == 2) {  // PDE filter then clamp, Filter.cc:73-102
        TOB_CHECK(pdefilter_apply(h, (const double*)dx, (double*)dxt));
        const int nbk = flat_grid(c, ne);
        k_clamp01<<<nbk, kFlatThreads, 0, c->stream>>>((double*)dxt, ne, c->d_partials);
        c->launches++;
        TOB_KERNEL_CHECK();
        TOB_CHECK(finalize_max(c, nbk, 1, F_S_TMP, true));
        double worst = 0;
        TOB_CHECK(read_scalars(c, F_S_TMP, 1, &worst));
        if (worst > 1.0e-4 && c->desc.rank == 0)
            printf("BOUND VIOLATION IN PDEFILTER - INCREASE RMIN OR MESH RESOLUTION: violation = %f\n", worst);
    } else {  // VecCopy(x, xTilde), Filter.cc:104-107
        if (dxt != dx) TOB_CHECK(vec_copy(c, (double*)dxt, (const double*)dx, ne));
    }
    if (projection) {  // HeavisideFilter(xPhys, xTilde), Filter.cc:110-111
        k_heaviside<<<flat_grid(c, ne), kFlatThreads, 0, c->stream>>>((const double*)dxt, (double*)dxp, ne, beta, eta);
        c->launches++;
        TOB_KERNEL_CHECK();
    } else {
        TOB_CHECK(vec_copy(c, (double*)dxp, (const double*)dxt, ne));
    }
    TOB_CHECK(stage_out_finish(c, xTilde, dxt, nb));
    return stage_out_finish(c, xPhys, dxp, nb);
}

int tob_filter_gradients(tob_filter* hh, const double* x, const double* xTilde, double* dfdx, int m, double** dgdx,
                         int projection, double beta, double eta) {
    auto* h = reinterpret_cast<tob_filter_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const long long ne = c->g0.owned_elems();
    const size_t nb = sizeof(double) * ne;
    if (m < 0 || m > 8) TOB_FAIL("at most 8 constraints supported");
    // in/out gradient vectors: stage host ones into private device buffers (slots are too few for 1+m)
    std::vector<double*> dev(1 + m, nullptr);
    std::vector<double*> host(1 + m, nullptr);
    host[0] = dfdx;
    for (int i = 0; i < m; i++) host[1 + i] = dgdx[i];
    if ((int)h->gbuf.size() < 1 + m) h->gbuf.resize(1 + m);
    for (int k = 0; k < 1 + m; k++) {
        if (!host[k]) TOB_FAIL("null gradient vector");
        if (is_device_ptr(host[k])) dev[k] = host[k];
        else {
            TOB_CHECK(h->gbuf[k].reserve(nb));
            dev[k] = h->gbuf[k].as<double>();
            TOB_CUDA(cudaMemcpyAsync(dev[k], host[k], nb, cudaMemcpyHostToDevice, c->stream));
        }
    }
    const void* dx; const void* dxt;
    TOB_CHECK(stage_in(c, 0, x, nb, &dx));
    TOB_CHECK(stage_in(c, 1, xTilde, nb, &dxt));
    if (projection) {  // chain rule, Filter.cc:125-164
        GradPtrs gp;
        gp.n = 1 + m;
        for (int k = 0; k < 1 + m; k++) gp.p[k] = dev[k];
        k_heaviside_chain<<<flat_grid(c, ne), kFlatThreads, 0, c->stream>>>((const double*)dxt, gp, ne, beta, eta);
        c->launches++;
        TOB_KERNEL_CHECK();
    }
    if (h->type == 0) {  // dfdx = H(dfdx.x)/Hs/x, Filter.cc:167-177 (objective only)
        TOB_CHECK(to_filter(h, dev[0], (const double*)dx, h->f_st, 2));
        TOB_CHECK(cone_apply(h, h->f_st, dev[0], (const double*)dx, 2));
    } else if (h->type == 1) {  // g = H(g/Hs), Filter.cc:178-192
        for (int k = 0; k < 1 + m; k++) {
            TOB_CHECK(to_filter(h, dev[k], nullptr, h->f_st, 1));
            TOB_CHECK(cone_apply(h, h->f_st, dev[k], nullptr, 1));
        }
    } else if (h->type == 2) {  // pdef->Gradients, Filter.cc:193-200
        for (int k = 0; k < 1 + m; k++) {
            TOB_CHECK(pdefilter_apply(h, dev[k], h->tmp_nat));
            TOB_CHECK(vec_copy(c, dev[k], h->tmp_nat, ne));
        }
    }
    bool any_host = false;
    for (int k = 0; k < 1 + m; k++)
        if (dev[k] != host[k]) {
            TOB_CUDA(cudaMemcpyAsync(host[k], dev[k], nb, cudaMemcpyDeviceToHost, c->stream));
            any_host = true;
        }
    if (any_host) TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int tob_filter_mnd(tob_filter* hh, const double* xPhys, double* mnd) {
    auto* h = reinterpret_cast<tob_filter_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = c->g0;
    const long long ne = g.owned_elems();
    const void* dx;
    TOB_CHECK(stage_in(c, 0, xPhys, sizeof(double) * ne, &dx));
    const int nbk = flat_grid(c, ne);
    k_mnd<<<nbk, kFlatThreads, 0, c->stream>>>((const double*)dx, ne, c->d_partials);
    c->launches++;
    TOB_KERNEL_CHECK();
    TOB_CHECK(finalize_sum(c, nbk, 1, F_S_MND, true));
    double s = 0;
    TOB_CHECK(read_scalars(c, F_S_MND, 1, &s));
    *mnd = s / ((double)g.nex * g.ney * g.nez);  // Filter.cc:221-222
    return 0;
}

int tob_filter_increase_beta(double* beta, double beta_final, double gx, int itr, double ch, int* changed) {
    // Filter::IncreaseBeta, Filter.cc:268-288
    int change = 0;
    if ((ch < 0.01 || itr % 10 == 0) && beta[0] < beta_final && gx < 0.000001) {
        change = 1;
        if (beta[0] < 7) beta[0] = beta[0] + 1;
        else beta[0] = beta[0] * 1.2;
        if (beta[0] > beta_final) { beta[0] = beta_final; change = 0; }
    }
    if (changed) *changed = change;
    return 0;
}

int tob_filter_get_hs(tob_filter* hh, double* hs) {
    auto* h = reinterpret_cast<tob_filter_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = c->g0;
    if (h->type != 0 && h->type != 1) TOB_FAIL("Hs exists only for filter types 0/1");
    const long long ne = g.owned_elems();
    void* d;
    TOB_CHECK(stage_out_begin(c, 1, hs, sizeof(double) * ne, &d));
    // stored_to_nat_elem expects the physics layout; the filter layout differs only by its ghost offset
    LevelGrid fg = g;
    fg.ge0 = h->gl0;
    TOB_CHECK(stored_to_nat_elem(c, fg, h->hs_st, (double*)d));
    return stage_out_finish(c, hs, d, sizeof(double) * ne);
}

int tob_filter_pde_apply(tob_filter* hh, const double* x, double* out) {
    auto* h = reinterpret_cast<tob_filter_impl*>(hh);
    tob_ctx_impl* c = h->c;
    if (h->type != 2) TOB_FAIL("not a PDE filter");
    const size_t nb = sizeof(double) * c->g0.owned_elems();
    const void* dx; void* dout;
    TOB_CHECK(stage_in(c, 0, x, nb, &dx));
    TOB_CHECK(stage_out_begin(c, 1, out, nb, &dout));
    TOB_CHECK(pdefilter_apply(h, (const double*)dx, (double*)dout));
    return stage_out_finish(c, out, dout, nb);
}

int tob_filter_last_stats(tob_filter* hh, tob_solve_stats* s) {
    *s = reinterpret_cast<tob_filter_impl*>(hh)->stats;
    return 0;
}

}  // extern "C"
