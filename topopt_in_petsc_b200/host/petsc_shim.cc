// petsc_shim.cc -- implementation of the petsc.h / mpi.h stand-ins (host side of the drop-in boundary).
// Vec = host array (pinned when CUDA is linked in), DM = grid handle + reference-counted tob_ctx,
// mini-MPI = POSIX shared memory between the per-GPU processes started by host/tobrun.
#include <fcntl.h>
#include <stdarg.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <atomic>
#include <map>
#include <vector>
#include "petsc.h"
#ifndef TOB_SHIM_HOST_ONLY
#include <cuda_runtime.h>
#include "../../include/topopt_b200.h"
#endif

// ---------------------------------------------------------------------------------------------
// mini-MPI
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kMaxRanks = 16;
constexpr int kMaxCount = 4096;
struct Shm {
    std::atomic<int> arrived;
    std::atomic<int> generation;
    std::atomic<int> id_ready;
    unsigned char nccl_id[128];
    double buf[kMaxRanks][kMaxCount];
};
int g_rank = 0, g_size = 1;
Shm* g_shm = nullptr;
std::string g_shm_name;
std::map<std::string, std::string> g_opts;

void shm_barrier() {
    if (g_size == 1) return;
    const int gen = g_shm->generation.load();
    if (g_shm->arrived.fetch_add(1) + 1 == g_size) {
        g_shm->arrived.store(0);
        g_shm->generation.fetch_add(1);
    } else {
        while (g_shm->generation.load() == gen) usleep(0);
    }
}
size_t type_size(MPI_Datatype t) { return t == MPI_DOUBLE ? 8 : (t == MPI_INT ? 4 : 8); }
}  // namespace

int TobRank(void) { return g_rank; }
int TobSize(void) { return g_size; }

extern "C" {
int MPI_Init(int*, char***) {
    const char* r = getenv("TOB_RANK");
    const char* n = getenv("TOB_NRANKS");
    g_rank = r ? atoi(r) : 0;
    g_size = n ? atoi(n) : 1;
    if (g_size > kMaxRanks) { fprintf(stderr, "tob mini-MPI: at most %d ranks\n", kMaxRanks); exit(2); }
    if (g_size > 1) {
        const char* job = getenv("TOB_JOB");
        g_shm_name = std::string("/tob_shm_") + (job ? job : "default");
        int fd = shm_open(g_shm_name.c_str(), O_CREAT | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, sizeof(Shm)) != 0) { perror("tob mini-MPI shm"); exit(2); }
        g_shm = (Shm*)mmap(nullptr, sizeof(Shm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (g_shm == MAP_FAILED) { perror("tob mini-MPI mmap"); exit(2); }
    }
    return 0;
}
int MPI_Finalize(void) {
    if (g_shm) {
        shm_barrier();
        munmap(g_shm, sizeof(Shm));
        if (g_rank == 0) shm_unlink(g_shm_name.c_str());
        g_shm = nullptr;
    }
    return 0;
}
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = g_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = g_size; return 0; }
int MPI_Barrier(MPI_Comm) { shm_barrier(); return 0; }
int MPI_Abort(MPI_Comm, int code) { exit(code ? code : 1); }
double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
int MPI_Allreduce(const void* sb, void* rb, int count, MPI_Datatype type, MPI_Op op, MPI_Comm) {
    const size_t ts = type_size(type);
    if (g_size == 1) { if (sb != rb) memcpy(rb, sb, ts * count); return 0; }
    if ((size_t)count * ts > sizeof(double) * kMaxCount) { fprintf(stderr, "tob mini-MPI: Allreduce too large\n"); exit(2); }
    memcpy(g_shm->buf[g_rank], sb, ts * count);
    shm_barrier();
    for (int i = 0; i < count; i++) {
        if (type == MPI_DOUBLE) {
            double acc = ((double*)g_shm->buf[0])[i];
            for (int r = 1; r < g_size; r++) {
                const double v = ((double*)g_shm->buf[r])[i];
                acc = op == MPI_SUM ? acc + v : (op == MPI_MAX ? (v > acc ? v : acc) : (v < acc ? v : acc));
            }
            ((double*)rb)[i] = acc;
        } else if (type == MPI_INT) {
            int acc = ((int*)g_shm->buf[0])[i];
            for (int r = 1; r < g_size; r++) {
                const int v = ((int*)g_shm->buf[r])[i];
                acc = op == MPI_SUM ? acc + v : (op == MPI_MAX ? (v > acc ? v : acc) : (v < acc ? v : acc));
            }
            ((int*)rb)[i] = acc;
        } else {
            unsigned long acc = ((unsigned long*)g_shm->buf[0])[i];
            for (int r = 1; r < g_size; r++) {
                const unsigned long v = ((unsigned long*)g_shm->buf[r])[i];
                acc = op == MPI_SUM ? acc + v : (op == MPI_MAX ? (v > acc ? v : acc) : (v < acc ? v : acc));
            }
            ((unsigned long*)rb)[i] = acc;
        }
    }
    shm_barrier();
    return 0;
}
int MPI_Allgather(const void* sb, int scount, MPI_Datatype st, void* rb, int, MPI_Datatype, MPI_Comm) {
    const size_t nb = type_size(st) * scount;
    if (g_size == 1) { memcpy(rb, sb, nb); return 0; }
    memcpy(g_shm->buf[g_rank], sb, nb);
    shm_barrier();
    for (int r = 0; r < g_size; r++) memcpy((char*)rb + r * nb, g_shm->buf[r], nb);
    shm_barrier();
    return 0;
}
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm) {
    if (g_size == 1) return 0;
    const size_t nb = type_size(type) * count;
    if (g_rank == root) memcpy(g_shm->buf[root], buf, nb);
    shm_barrier();
    if (g_rank != root) memcpy(buf, g_shm->buf[root], nb);
    shm_barrier();
    return 0;
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------
// init / options / printing
// ---------------------------------------------------------------------------------------------
PetscErrorCode PetscInitialize(int* argc, char*** argv, const char*, const char*) {
    MPI_Init(argc, argv);
    g_opts.clear();
    for (int i = 1; argc && i < *argc; i++) {
        const char* a = (*argv)[i];
        if (a[0] != '-') continue;
        std::string key(a), val = "1";
        if (i + 1 < *argc) {
            const char* nx = (*argv)[i + 1];
            const bool numeric_neg = nx[0] == '-' && (isdigit((unsigned char)nx[1]) || nx[1] == '.');
            if (nx[0] != '-' || numeric_neg) { val = nx; i++; }
        }
        g_opts[key] = val;
    }
    return 0;
}
PetscErrorCode PetscFinalize(void) { return MPI_Finalize(); }
PetscErrorCode PetscPrintf(MPI_Comm, const char* fmt, ...) {
    if (g_rank != 0) return 0;
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    fflush(stdout);
    return 0;
}
PetscErrorCode PetscErrorPrintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    return 0;
}
static const std::string* find_opt(const char* name) {
    auto it = g_opts.find(name);
    return it == g_opts.end() ? nullptr : &it->second;
}
PetscErrorCode PetscOptionsGetInt(void*, const char*, const char* name, PetscInt* v, PetscBool* set) {
    const std::string* s = find_opt(name);
    if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
    if (s) *v = atoi(s->c_str());
    return 0;
}
PetscErrorCode PetscOptionsGetReal(void*, const char*, const char* name, PetscReal* v, PetscBool* set) {
    const std::string* s = find_opt(name);
    if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
    if (s) *v = atof(s->c_str());
    return 0;
}
PetscErrorCode PetscOptionsGetBool(void*, const char*, const char* name, PetscBool* v, PetscBool* set) {
    const std::string* s = find_opt(name);
    if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
    if (s) {
        const std::string& t = *s;
        *v = (t == "0" || t == "false" || t == "no" || t == "FALSE" || t == "PETSC_FALSE") ? PETSC_FALSE : PETSC_TRUE;
    }
    return 0;
}
PetscErrorCode PetscOptionsGetString(void*, const char*, const char* name, char* v, size_t len, PetscBool* set) {
    const std::string* s = find_opt(name);
    if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
    if (s) { strncpy(v, s->c_str(), len - 1); v[len - 1] = 0; }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Vec
// ---------------------------------------------------------------------------------------------
struct _p_Vec {
    PetscInt nloc = 0, nglob = 0;
    PetscScalar* a = nullptr;
};
static PetscScalar* host_alloc(size_t n) {
    PetscScalar* p = nullptr;
#ifndef TOB_SHIM_HOST_ONLY
    if (cudaMallocHost((void**)&p, sizeof(PetscScalar) * (n ? n : 1)) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
    if (p) { memset(p, 0, sizeof(PetscScalar) * n); return p; }
#endif
    return (PetscScalar*)calloc(n ? n : 1, sizeof(PetscScalar));
}
static void host_free(PetscScalar* p) {
#ifndef TOB_SHIM_HOST_ONLY
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost) { cudaFreeHost(p); return; }
    cudaGetLastError();
#endif
    free(p);
}
PetscErrorCode VecCreateTob(PetscInt nlocal, Vec* v) {
    Vec x = new _p_Vec();
    x->nloc = nlocal;
    int tot = nlocal;
    MPI_Allreduce(&nlocal, &tot, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    x->nglob = tot;
    x->a = host_alloc(nlocal);
    *v = x;
    return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec* out) {
    Vec x = new _p_Vec();
    x->nloc = v->nloc; x->nglob = v->nglob;
    x->a = host_alloc(v->nloc);
    *out = x;
    return 0;
}
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt m, Vec** out) {
    Vec* arr = new Vec[m];
    for (PetscInt i = 0; i < m; i++) VecDuplicate(v, &arr[i]);
    *out = arr;
    return 0;
}
PetscErrorCode VecDestroy(Vec* v) {
    if (v && *v) { host_free((*v)->a); delete *v; *v = nullptr; }
    return 0;
}
PetscErrorCode VecDestroyVecs(PetscInt m, Vec** v) {
    if (v && *v) { for (PetscInt i = 0; i < m; i++) VecDestroy(&(*v)[i]); delete[] *v; *v = nullptr; }
    return 0;
}
PetscErrorCode VecGetArray(Vec v, PetscScalar** a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec, PetscScalar** a) { if (a) *a = nullptr; return 0; }
PetscErrorCode VecGetArrays(const Vec* v, PetscInt m, PetscScalar*** a) {
    PetscScalar** p = new PetscScalar*[m];
    for (PetscInt i = 0; i < m; i++) p[i] = v[i]->a;
    *a = p;
    return 0;
}
PetscErrorCode VecRestoreArrays(const Vec*, PetscInt, PetscScalar*** a) {
    if (a && *a) { delete[] *a; *a = nullptr; }
    return 0;
}
PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n) { *n = v->nloc; return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt* n) { *n = v->nglob; return 0; }
PetscErrorCode VecSet(Vec v, PetscScalar a) { for (PetscInt i = 0; i < v->nloc; i++) v->a[i] = a; return 0; }
PetscErrorCode VecCopy(Vec x, Vec y) { memcpy(y->a, x->a, sizeof(PetscScalar) * x->nloc); return 0; }
PetscErrorCode VecScale(Vec v, PetscScalar a) { for (PetscInt i = 0; i < v->nloc; i++) v->a[i] *= a; return 0; }
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) { for (PetscInt i = 0; i < y->nloc; i++) y->a[i] += a * x->a[i]; return 0; }
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar a, PetscScalar b, PetscScalar g, Vec x, Vec y) {
    for (PetscInt i = 0; i < z->nloc; i++) z->a[i] = a * x->a[i] + b * y->a[i] + g * z->a[i];
    return 0;
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y) { for (PetscInt i = 0; i < w->nloc; i++) w->a[i] = x->a[i] * y->a[i]; return 0; }
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y) { for (PetscInt i = 0; i < w->nloc; i++) w->a[i] = x->a[i] / y->a[i]; return 0; }
PetscErrorCode VecSum(Vec v, PetscScalar* s) {
    double loc = 0, tot = 0;
    for (PetscInt i = 0; i < v->nloc; i++) loc += v->a[i];
    MPI_Allreduce(&loc, &tot, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    *s = tot;
    return 0;
}
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal* n) {
    double loc = 0, tot = 0;
    if (t == NORM_INFINITY) {
        for (PetscInt i = 0; i < v->nloc; i++) loc = fmax(loc, fabs(v->a[i]));
        MPI_Allreduce(&loc, &tot, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
        *n = tot;
    } else if (t == NORM_1) {
        for (PetscInt i = 0; i < v->nloc; i++) loc += fabs(v->a[i]);
        MPI_Allreduce(&loc, &tot, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
        *n = tot;
    } else {
        for (PetscInt i = 0; i < v->nloc; i++) loc += v->a[i] * v->a[i];
        MPI_Allreduce(&loc, &tot, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
        *n = sqrt(tot);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// viewers.  Binary Vec record of PETSc's VecView/VecLoad: int32 VEC_FILE_CLASSID (1211214), int32 global
// length, then the entries as big-endian IEEE doubles in global (= natural, for z-slabs) order.
// ---------------------------------------------------------------------------------------------
struct _p_PetscViewer {
    FILE* f = nullptr;
    bool binary = true;
    PetscFileMode mode = FILE_MODE_READ;
    std::string name;
};
static void swap_copy(void* dst, const void* src, size_t elem, size_t n) {
    const unsigned char* s = (const unsigned char*)src;
    unsigned char* d = (unsigned char*)dst;
    for (size_t i = 0; i < n; i++)
        for (size_t b = 0; b < elem; b++) d[i * elem + b] = s[i * elem + elem - 1 - b];
}
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm, const char* name, PetscFileMode mode, PetscViewer* v) {
    PetscViewer w = new _p_PetscViewer();
    w->binary = true; w->mode = mode; w->name = name;
    if (mode == FILE_MODE_WRITE) {
        if (g_rank == 0) { w->f = fopen(name, "wb"); if (w->f) fclose(w->f); }  // truncate once
        MPI_Barrier(MPI_COMM_WORLD);
        w->f = nullptr;
    } else {
        w->f = fopen(name, "rb");
        if (!w->f) { delete w; *v = nullptr; return 66; }
    }
    *v = w;
    return 0;
}
PetscErrorCode PetscViewerASCIIOpen(MPI_Comm, const char* name, PetscViewer* v) {
    PetscViewer w = new _p_PetscViewer();
    w->binary = false; w->mode = FILE_MODE_WRITE; w->name = name;
    if (g_rank == 0) w->f = fopen(name, "w");
    *v = w;
    return 0;
}
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char* fmt, ...) {
    if (!v || !v->f) return 0;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(v->f, fmt, ap);
    va_end(ap);
    return 0;
}
PetscErrorCode PetscViewerDestroy(PetscViewer* v) {
    if (v && *v) { if ((*v)->f) fclose((*v)->f); delete *v; *v = nullptr; }
    return 0;
}
PetscErrorCode VecView(Vec x, PetscViewer v) {
    if (!v) return 66;
    // ranks append their slabs in rank order (natural order == PETSc global order for z-slabs)
    for (int r = 0; r < g_size; r++) {
        if (r == g_rank) {
            FILE* f = fopen(v->name.c_str(), "ab");
            if (!f) return 66;
            if (r == 0) {
                int hdr[2] = {1211214, x->nglob}, be[2];
                swap_copy(be, hdr, 4, 2);
                fwrite(be, 4, 2, f);
            }
            std::vector<double> be(x->nloc);
            swap_copy(be.data(), x->a, 8, x->nloc);
            fwrite(be.data(), 8, x->nloc, f);
            fclose(f);
        }
        MPI_Barrier(MPI_COMM_WORLD);
    }
    return 0;
}
PetscErrorCode VecLoad(Vec x, PetscViewer v) {
    if (!v || !v->f) return 66;
    int be[2], hdr[2];
    if (fread(be, 4, 2, v->f) != 2) return 66;
    swap_copy(hdr, be, 4, 2);
    if (hdr[0] != 1211214 || hdr[1] != x->nglob) return 66;
    // every rank reads the whole record and keeps its slab
    std::vector<int> counts(g_size);
    int n = x->nloc;
    MPI_Allgather(&n, 1, MPI_INT, counts.data(), 1, MPI_INT, MPI_COMM_WORLD);
    long off = 0;
    for (int r = 0; r < g_rank; r++) off += counts[r];
    std::vector<double> all(x->nglob);
    if (fread(all.data(), 8, x->nglob, v->f) != (size_t)x->nglob) return 66;
    swap_copy(x->a, all.data() + off, 8, x->nloc);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// DM
// ---------------------------------------------------------------------------------------------
struct CtxHolder {
    tob_ctx* ctx = nullptr;
    int refs = 0;
};
static tob_ctx* g_default_ctx = nullptr;  // the grid TopOpt built; the device MMA flavour attaches to it
tob_ctx* TobDefaultCtx() { return g_default_ctx; }
struct _p_DM {
    PetscInt n[3] = {0, 0, 0};
    PetscScalar xc[6] = {0, 0, 0, 0, 0, 0};
    PetscInt dof = 1;
    CtxHolder* holder = nullptr;
};
PetscErrorCode DMCreateTob(PetscInt nxn, PetscInt nyn, PetscInt nzn, const PetscScalar xc[6], PetscInt dof, DM* dm) {
    DM d = new _p_DM();
    d->n[0] = nxn; d->n[1] = nyn; d->n[2] = nzn; d->dof = dof;
    memcpy(d->xc, xc, sizeof(PetscScalar) * 6);
    d->holder = new CtxHolder();
    d->holder->refs = 1;
#ifndef TOB_SHIM_HOST_ONLY
    tob_grid_desc gd;
    memset(&gd, 0, sizeof gd);
    gd.nxn = nxn; gd.nyn = nyn; gd.nzn = nzn;
    for (int i = 0; i < 6; i++) gd.xc[i] = xc[i];
    gd.rank = g_rank; gd.nranks = g_size; gd.device = -1; gd.stream = nullptr;
    unsigned char id[128];
    if (g_size > 1) {  // rank 0 creates the ncclUniqueId, the others read it from the shared segment
        if (g_rank == 0) {
            if (tob_nccl_unique_id(id) != 0) { fprintf(stderr, "%s\n", tob_last_error()); return 77; }
            memcpy(g_shm->nccl_id, id, 128);
        }
        MPI_Barrier(MPI_COMM_WORLD);
        memcpy(id, g_shm->nccl_id, 128);
        MPI_Barrier(MPI_COMM_WORLD);
        gd.nccl_id = id;
    }
    if (tob_ctx_create(&gd, &d->holder->ctx) != 0) {
        fprintf(stderr, "tob_ctx_create: %s\n", tob_last_error());
        return 77;
    }
    g_default_ctx = d->holder->ctx;
#endif
    *dm = d;
    return 0;
}
PetscErrorCode DMCloneTob(DM dm, PetscInt dof, DM* out) {
    DM d = new _p_DM(*dm);
    d->dof = dof;
    d->holder->refs++;
    *out = d;
    return 0;
}
PetscErrorCode DMDestroy(DM* dm) {
    if (!dm || !*dm) return 0;
    CtxHolder* h = (*dm)->holder;
    if (h && --h->refs == 0) {
#ifndef TOB_SHIM_HOST_ONLY
        if (h->ctx) tob_ctx_destroy(h->ctx);
#endif
        delete h;
    }
    delete *dm;
    *dm = nullptr;
    return 0;
}
tob_ctx* DMGetTobCtx(DM dm) { return dm->holder ? dm->holder->ctx : nullptr; }
PetscErrorCode DMGetGridTob(DM dm, PetscInt n[3], PetscScalar xc[6]) {
    for (int i = 0; i < 3; i++) n[i] = dm->n[i];
    for (int i = 0; i < 6; i++) xc[i] = dm->xc[i];
    return 0;
}
