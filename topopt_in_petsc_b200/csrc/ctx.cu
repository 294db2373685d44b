// ctx.cu -- context (grid, z-slab partition, stream, NCCL communicator) and host<->device staging.
// Replaces what the reference gets from DMDACreate3d/DMCreateGlobalVector (TopOpt.cc:233-300,
// LinearElasticity.cc:97-110): here the "DM" is index arithmetic on a structured slab.
#include "common.cuh"

namespace tob {

thread_local std::string g_last_error;

int set_error(const char* file, int line, const std::string& msg) {
    char buf[64];
    snprintf(buf, sizeof buf, ":%d: ", line);
    const char* base = strrchr(file, '/');
    g_last_error = std::string(base ? base + 1 : file) + buf + msg;
    return 1;
}

static int round_up(int a, int m) { return (a + m - 1) / m * m; }

void make_level_grid(LevelGrid& g, int nxn, int nyn, int nzn, double dx, double dy, double dz, int rank, int nranks,
                     int kn0, int kn1) {
    g.nxn = nxn; g.nyn = nyn; g.nzn = nzn;
    g.nex = nxn - 1; g.ney = nyn - 1; g.nez = nzn - 1;
    g.dx = dx; g.dy = dy; g.dz = dz;
    g.kn0 = kn0; g.kn1 = kn1;
    // an element belongs to the rank owning its upper (k+1) node plane (LinearElasticity.cc:804-818)
    g.ke0 = (rank == 0) ? 0 : kn0 - 1;
    g.ke1 = kn1 - 1;
    g.gn0 = kn0 > 0 ? kn0 - 1 : 0;
    g.gn1 = kn1 < nzn ? kn1 + 1 : nzn;
    g.ge0 = g.gn0;
    g.ge1 = g.gn1 - 1;
    g.px = round_up(nxn, 4);
    g.pe = round_up(g.nex, 4);
    g.pm = round_up(nxn, 16);
    g.mstride = (long long)g.pm * nyn;
    g.nstride = (long long)g.px * nyn;
    g.pstride = 3 * g.nstride;
    g.estride = (long long)g.pe * g.ney;
    g.nvec = g.pstride * g.nzl();
    g.evec = g.estride * g.nzle();
    g.own_off = g.pstride * (kn0 - g.gn0);
    g.own_len = g.pstride * (kn1 - kn0);
    (void)nranks;
}

int DevBuf::reserve(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 256;
    TOB_CUDA(cudaMalloc(&p, want));
    cap = want;
    return 0;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int stage_in(tob_ctx_impl* c, int slot, const void* src, size_t nbytes, const void** dev) {
    if (!src) TOB_FAIL("null input vector");
    if (is_device_ptr(src)) { *dev = src; return 0; }
    TOB_CHECK(c->stage[slot].reserve(nbytes));
    TOB_CUDA(cudaMemcpyAsync(c->stage[slot].p, src, nbytes, cudaMemcpyHostToDevice, c->stream));
    *dev = c->stage[slot].p;
    return 0;
}
int stage_out_begin(tob_ctx_impl* c, int slot, void* dst, size_t nbytes, void** dev) {
    if (!dst) TOB_FAIL("null output vector");
    if (is_device_ptr(dst)) { *dev = dst; return 0; }
    TOB_CHECK(c->stage[slot].reserve(nbytes));
    *dev = c->stage[slot].p;
    return 0;
}
int stage_out_finish(tob_ctx_impl* c, void* dst, void* dev, size_t nbytes) {
    if (dev == dst) return 0;
    TOB_CUDA(cudaMemcpyAsync(dst, dev, nbytes, cudaMemcpyDeviceToHost, c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int read_scalars(tob_ctx_impl* c, int slot0, int n, double* out) {
    TOB_CUDA(cudaMemcpyAsync(c->h_scalars + slot0, c->d_scalars + slot0, sizeof(double) * n, cudaMemcpyDeviceToHost,
                             c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) out[i] = c->h_scalars[slot0 + i];
    return 0;
}

}  // namespace tob

using namespace tob;

extern "C" {

const char* tob_last_error(void) { return g_last_error.c_str(); }
int tob_version(void) { return 100; }

int tob_ctx_create(const tob_grid_desc* d, tob_ctx** out) {
    if (!d || !out) TOB_FAIL("null argument");
    if (d->nxn < 2 || d->nyn < 2 || d->nzn < 2) TOB_FAIL("grid needs at least 2 nodes per direction");
    if (d->nranks < 1 || d->rank < 0 || d->rank >= d->nranks) TOB_FAIL("bad rank/nranks");
    if (d->nzn - 1 < d->nranks) TOB_FAIL("fewer element layers than z-slabs");
    int ndev = 0;
    TOB_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev == 0) TOB_FAIL("no CUDA device: libtopopt_b200 has no CPU fallback");
    auto* c = new tob_ctx_impl();
    c->desc = *d;
    c->device = d->device >= 0 ? d->device : d->rank % ndev;
    TOB_CUDA(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    TOB_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 9) TOB_FAIL("libtopopt_b200 needs TMA/mbarrier hardware (built for sm_100a)");
    c->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("TOB_L2HINT")) c->l2hint = atoi(e);
    if (c->l2hint & 8) {
        // evict_last lines are only protected inside the persisting set-aside of the L2 (default size 0)
        size_t want = (size_t)prop.persistingL2CacheMaxSize;
        if (const char* e = getenv("TOB_L2_PERSIST_MB")) want = std::min(want, (size_t)atoi(e) << 20);
        TOB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        if (getenv("TOB_VERBOSE")) fprintf(stderr, "[tob] persisting L2 set-aside %zu MB (max %d MB, L2 %d MB)\n", want >> 20,
                                           prop.persistingL2CacheMaxSize >> 20, prop.l2CacheSize >> 20);
    }
    if (d->stream) { c->stream = (cudaStream_t)d->stream; c->own_stream = false; }
    else { TOB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    TOB_CUDA(cudaEventCreate(&c->ev0));
    TOB_CUDA(cudaEventCreate(&c->ev1));
    // z-slab partition: element layers split evenly, rank 0 additionally owns node plane 0
    int kn0 = 0, kn1 = 0;
    TOB_CHECK(tob_slab_partition(d->nzn, d->nranks, d->rank, &kn0, &kn1, nullptr, nullptr));
    const double dx = (d->xc[1] - d->xc[0]) / (d->nxn - 1);  // TopOpt.cc:225-227
    const double dy = (d->xc[3] - d->xc[2]) / (d->nyn - 1);
    const double dz = (d->xc[5] - d->xc[4]) / (d->nzn - 1);
    make_level_grid(c->g0, d->nxn, d->nyn, d->nzn, dx, dy, dz, d->rank, d->nranks, kn0, kn1);
    c->n_partials = kMaxPartials;
    TOB_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * c->n_partials));
    TOB_CUDA(cudaMalloc(&c->d_scalars, sizeof(double) * kScalarSlots));
    TOB_CUDA(cudaMemsetAsync(c->d_scalars, 0, sizeof(double) * kScalarSlots, c->stream));
    TOB_CUDA(cudaMallocHost(&c->h_scalars, sizeof(double) * kScalarSlots));
    cudaDriverEntryPointQueryResult qres;
    TOB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->encode_tiled, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !c->encode_tiled) TOB_FAIL("cuTensorMapEncodeTiled unavailable");
    if (d->nranks > 1) {
        if (!d->nccl_id) TOB_FAIL("nranks>1 needs a shared ncclUniqueId");
        ncclUniqueId id;
        memcpy(&id, d->nccl_id, sizeof id);
        TOB_NCCL(ncclCommInitRank(&c->comm, d->nranks, id, d->rank));
        TOB_CHECK(peer_setup(c, sizeof(double) * (size_t)c->g0.pstride));
    }
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    *out = reinterpret_cast<tob_ctx*>(c);
    return 0;
}

int tob_slab_partition(int nzn, int nranks, int rank, int* kn0, int* kn1, int* ke0, int* ke1) {
    if (nzn < 2 || nranks < 1 || rank < 0 || rank >= nranks) TOB_FAIL("bad slab partition arguments");
    if (nzn - 1 < nranks) TOB_FAIL("fewer element layers than z-slabs");
    const int nez = nzn - 1;
    const int es = (int)((long long)nez * rank / nranks), ee = (int)((long long)nez * (rank + 1) / nranks);
    const int n0 = rank == 0 ? 0 : es + 1, n1 = ee + 1;
    if (kn0) *kn0 = n0;
    if (kn1) *kn1 = n1;
    if (ke0) *ke0 = es;   // == (rank == 0 ? 0 : n0 - 1): elements follow their upper node plane
    if (ke1) *ke1 = ee;
    return 0;
}

int tob_ctx_destroy(tob_ctx* h) {
    if (!h) return 0;
    auto* c = reinterpret_cast<tob_ctx_impl*>(h);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    peer_teardown(c);
    if (c->comm) ncclCommDestroy(c->comm);
    for (auto& b : c->stage) b.release();
    cudaFree(c->d_partials);
    cudaFree(c->d_scalars);
    cudaFreeHost(c->h_scalars);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int tob_ctx_sizes(tob_ctx* h, long long* nel_owned, long long* ndof_owned, long long* nel_global,
                  long long* ndof_global) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(h);
    const LevelGrid& g = c->g0;
    if (nel_owned) *nel_owned = g.owned_elems();
    if (ndof_owned) *ndof_owned = 3 * g.owned_nodes();
    if (nel_global) *nel_global = (long long)g.nex * g.ney * g.nez;
    if (ndof_global) *ndof_global = 3LL * g.nxn * g.nyn * g.nzn;
    return 0;
}

int tob_ctx_slab(tob_ctx* h, int* kn0, int* kn1, int* ke0, int* ke1) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(h);
    if (kn0) *kn0 = c->g0.kn0;
    if (kn1) *kn1 = c->g0.kn1;
    if (ke0) *ke0 = c->g0.ke0;
    if (ke1) *ke1 = c->g0.ke1;
    return 0;
}

int tob_ctx_sync(tob_ctx* h) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(h);
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

long long tob_ctx_launch_count(tob_ctx* h, int reset) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(h);
    long long n = c->launches;
    if (reset) c->launches = 0;
    return n;
}

void* tob_ctx_stream(tob_ctx* h) { return reinterpret_cast<tob_ctx_impl*>(h)->stream; }

int tob_dev_alloc(tob_ctx* h, long long nbytes, void** out) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(h);
    TOB_CUDA(cudaSetDevice(c->device));
    TOB_CUDA(cudaMalloc(out, (size_t)nbytes));
    TOB_CUDA(cudaMemsetAsync(*out, 0, (size_t)nbytes, c->stream));
    return 0;
}
int tob_dev_free(tob_ctx* h, void* p) {
    (void)h;
    TOB_CUDA(cudaFree(p));
    return 0;
}
int tob_nccl_unique_id(void* out128) {
    ncclUniqueId id;
    TOB_NCCL(ncclGetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    memcpy(out128, &id, sizeof id);
    return 0;
}
int tob_memcpy(tob_ctx* h, void* dst, const void* src, long long nbytes) {
    auto* c = reinterpret_cast<tob_ctx_impl*>(h);
    TOB_CUDA(cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDefault, c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
