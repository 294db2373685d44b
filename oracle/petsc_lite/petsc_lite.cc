// petsc_lite.cc -- serial CPU implementation of oracle/petsc_lite/petsc.h + mpi.h (see petsc.h for scope).
// TEST INFRASTRUCTURE: lets the reference's own unmodified sources run on one process so that their outputs pin
// the oracles and the CUDA product.  Nothing here is linked into the product.
#include <petsc.h>
#include <stdarg.h>
#include <stdint.h>
#include <time.h>
#include <unistd.h>
#include <map>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct _p_Vec {
    PetscInt n;
    PetscScalar* a;
};
struct _p_Mat {
    PetscInt m, n;
    std::vector<std::map<PetscInt, PetscScalar>> rows;  // assembly form
    // CSR snapshot for the products (rebuilt lazily after any modification)
    bool csr_valid = false;
    std::vector<PetscInt> rp, ci;
    std::vector<PetscScalar> va;
    void csr() {
        if (csr_valid) return;
        rp.assign((size_t)m + 1, 0);
        size_t nnz = 0;
        for (PetscInt i = 0; i < m; i++) { nnz += rows[(size_t)i].size(); rp[(size_t)i + 1] = (PetscInt)nnz; }
        ci.resize(nnz);
        va.resize(nnz);
        size_t k = 0;
        for (PetscInt i = 0; i < m; i++)
            for (const auto& kv : rows[(size_t)i]) { ci[k] = kv.first; va[k] = kv.second; k++; }
        csr_valid = true;
    }
    void apply(const PetscScalar* x, PetscScalar* y) {
        csr();
#pragma omp parallel for schedule(static)
        for (PetscInt i = 0; i < m; i++) {
            PetscScalar s = 0;
            for (PetscInt k = rp[(size_t)i]; k < rp[(size_t)i + 1]; k++) s += va[(size_t)k] * x[ci[(size_t)k]];
            y[i] = s;
        }
    }
};
struct _p_PC {
    std::string type = "none";
    std::vector<KSP> sub;  // smoothers / coarse solve handed out by PCMGGet*
};
struct _p_KSP {
    std::string type = "gmres";
    Mat A = nullptr;
    PetscReal rtol = 1e-5, atol = 1e-50, dtol = 1e5;
    PetscInt maxits = 10000, its = 0, restart = 30;
    PetscReal rnorm = 0.0;
    PetscBool guess_nonzero = PETSC_FALSE;
    PC pc = nullptr;
};
struct _p_PetscViewer {
    FILE* f = nullptr;
    bool binary = false;
    PetscFileMode mode = FILE_MODE_WRITE;
    std::string name;
};
struct _p_PetscRandom {
    unsigned long long state = 0x1234ABCD330EULL;
};
struct _p_ISL2G {
    int dummy;
};
static _p_ISL2G g_l2g;

// ------------------------------------------------------------------------------------------------
// init / options / printing
// ------------------------------------------------------------------------------------------------
static std::vector<std::string> g_args;
PetscErrorCode PetscInitialize(int* argc, char*** argv, const char*, const char*) {
    g_args.clear();
    if (argc && argv)
        for (int i = 1; i < *argc; i++) g_args.push_back((*argv)[i]);
    return 0;
}
PetscErrorCode PetscFinalize(void) { return 0; }
PetscErrorCode PetscPrintf(MPI_Comm, const char* fmt, ...) {
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
static const char* find_opt(const char* name, bool* present) {
    *present = false;
    for (size_t i = 0; i < g_args.size(); i++)
        if (g_args[i] == name) {
            *present = true;
            if (i + 1 < g_args.size()) {
                const std::string& nx = g_args[i + 1];
                const bool is_flag = nx.size() > 1 && nx[0] == '-' && !(nx[1] >= '0' && nx[1] <= '9') && nx[1] != '.';
                if (!is_flag) return g_args[i + 1].c_str();
            }
            return nullptr;
        }
    return nullptr;
}
PetscErrorCode PetscOptionsGetInt(void*, const char*, const char* name, PetscInt* v, PetscBool* set) {
    bool p;
    const char* s = find_opt(name, &p);
    if (set) *set = (p && s) ? PETSC_TRUE : PETSC_FALSE;
    if (p && s) *v = (PetscInt)strtol(s, nullptr, 10);
    return 0;
}
PetscErrorCode PetscOptionsGetReal(void*, const char*, const char* name, PetscReal* v, PetscBool* set) {
    bool p;
    const char* s = find_opt(name, &p);
    if (set) *set = (p && s) ? PETSC_TRUE : PETSC_FALSE;
    if (p && s) *v = strtod(s, nullptr);
    return 0;
}
PetscErrorCode PetscOptionsGetBool(void*, const char*, const char* name, PetscBool* v, PetscBool* set) {
    bool p;
    const char* s = find_opt(name, &p);
    if (set) *set = p ? PETSC_TRUE : PETSC_FALSE;
    if (p) {
        if (!s) *v = PETSC_TRUE;
        else {
            const std::string t(s);
            *v = (t == "0" || t == "false" || t == "no" || t == "FALSE") ? PETSC_FALSE : PETSC_TRUE;
        }
    }
    return 0;
}
PetscErrorCode PetscOptionsGetString(void*, const char*, const char* name, char* v, size_t len, PetscBool* set) {
    bool p;
    const char* s = find_opt(name, &p);
    if (set) *set = (p && s) ? PETSC_TRUE : PETSC_FALSE;
    if (p && s) {
        strncpy(v, s, len - 1);
        v[len - 1] = 0;
    }
    return 0;
}
PetscErrorCode PetscObjectTypeCompare(PetscObject obj, const char* type, PetscBool* same) {
    // only ever asked about a PC (LinearElasticity.cc:680, PDEFilter.cc:299)
    *same = (obj && static_cast<PC>(obj)->type == type) ? PETSC_TRUE : PETSC_FALSE;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Vec
// ------------------------------------------------------------------------------------------------
PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec* v) {
    *v = new _p_Vec{n, (PetscScalar*)calloc((size_t)(n > 0 ? n : 1), sizeof(PetscScalar))};
    return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec* out) { return VecCreateSeq(0, v->n, out); }
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt m, Vec** out) {
    *out = (Vec*)malloc(sizeof(Vec) * (size_t)(m > 0 ? m : 1));
    for (PetscInt i = 0; i < m; i++) VecDuplicate(v, &(*out)[i]);
    return 0;
}
PetscErrorCode VecDestroy(Vec* v) {
    if (v && *v) {
        free((*v)->a);
        delete *v;
        *v = nullptr;
    }
    return 0;
}
PetscErrorCode VecDestroyVecs(PetscInt m, Vec** v) {
    if (v && *v) {
        for (PetscInt i = 0; i < m; i++) VecDestroy(&(*v)[i]);
        free(*v);
        *v = nullptr;
    }
    return 0;
}
PetscErrorCode VecGetArray(Vec v, PetscScalar** a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec, PetscScalar** a) { if (a) *a = nullptr; return 0; }
PetscErrorCode VecGetArrays(const Vec* v, PetscInt m, PetscScalar*** a) {
    *a = (PetscScalar**)malloc(sizeof(PetscScalar*) * (size_t)(m > 0 ? m : 1));
    for (PetscInt i = 0; i < m; i++) (*a)[i] = v[i]->a;
    return 0;
}
PetscErrorCode VecRestoreArrays(const Vec*, PetscInt, PetscScalar*** a) {
    if (a && *a) { free(*a); *a = nullptr; }
    return 0;
}
PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n) { *n = v->n; return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt* n) { *n = v->n; return 0; }
PetscErrorCode VecSet(Vec v, PetscScalar a) { for (PetscInt i = 0; i < v->n; i++) v->a[i] = a; return 0; }
PetscErrorCode VecSetValueLocal(Vec v, PetscInt i, PetscScalar a, InsertMode mode) {
    if (i < 0 || i >= v->n) return 63;
    if (mode == ADD_VALUES) v->a[i] += a; else v->a[i] = a;
    return 0;
}
PetscErrorCode VecAssemblyBegin(Vec) { return 0; }
PetscErrorCode VecAssemblyEnd(Vec) { return 0; }
PetscErrorCode VecCopy(Vec x, Vec y) { memcpy(y->a, x->a, sizeof(PetscScalar) * (size_t)x->n); return 0; }
PetscErrorCode VecScale(Vec v, PetscScalar a) { for (PetscInt i = 0; i < v->n; i++) v->a[i] *= a; return 0; }
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) { for (PetscInt i = 0; i < y->n; i++) y->a[i] += a * x->a[i]; return 0; }
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar a, PetscScalar b, PetscScalar g, Vec x, Vec y) {
    for (PetscInt i = 0; i < z->n; i++) z->a[i] = a * x->a[i] + b * y->a[i] + g * z->a[i];
    return 0;
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y) { for (PetscInt i = 0; i < w->n; i++) w->a[i] = x->a[i] * y->a[i]; return 0; }
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y) { for (PetscInt i = 0; i < w->n; i++) w->a[i] = x->a[i] / y->a[i]; return 0; }
PetscErrorCode VecSum(Vec v, PetscScalar* s) { PetscScalar t = 0; for (PetscInt i = 0; i < v->n; i++) t += v->a[i]; *s = t; return 0; }
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal* n) {
    PetscReal s = 0;
    for (PetscInt i = 0; i < v->n; i++) {
        const PetscReal a = fabs(v->a[i]);
        if (t == NORM_1) s += a; else if (t == NORM_2) s += a * a; else s = a > s ? a : s;
    }
    *n = t == NORM_2 ? sqrt(s) : s;
    return 0;
}
PetscErrorCode PetscRandomCreate(MPI_Comm, PetscRandom* r) { *r = new _p_PetscRandom(); return 0; }
PetscErrorCode PetscRandomSetType(PetscRandom, PetscRandomType) { return 0; }
PetscErrorCode PetscRandomDestroy(PetscRandom* r) { if (r && *r) { delete *r; *r = nullptr; } return 0; }
PetscErrorCode VecSetRandom(Vec v, PetscRandom r) {
    for (PetscInt i = 0; i < v->n; i++) {  // 48-bit LCG of drand48
        r->state = (0x5DEECE66DULL * r->state + 0xBULL) & ((1ULL << 48) - 1);
        v->a[i] = (double)r->state / (double)(1ULL << 48);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// viewers: PETSc binary Vec stream (big-endian: int32 classid 1211214, int32 n, n doubles), ASCII printer
// ------------------------------------------------------------------------------------------------
static uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
static uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm, const char* name, PetscFileMode mode, PetscViewer* v) {
    auto* p = new _p_PetscViewer();
    p->binary = true;
    p->mode = mode;
    p->name = name;
    p->f = fopen(name, mode == FILE_MODE_READ ? "rb" : "wb");
    *v = p;
    return p->f ? 0 : 65;
}
PetscErrorCode PetscViewerCreate(MPI_Comm, PetscViewer* v) { *v = new _p_PetscViewer(); return 0; }
PetscErrorCode PetscViewerSetType(PetscViewer v, PetscViewerType t) { v->binary = std::string(t) == PETSCVIEWERBINARY; return 0; }
PetscErrorCode PetscViewerFileSetMode(PetscViewer v, PetscFileMode mode) { v->mode = mode; return 0; }
PetscErrorCode PetscViewerFileSetName(PetscViewer v, const char* name) {
    if (v->f) fclose(v->f);
    v->name = name;
    v->f = fopen(name, v->mode == FILE_MODE_READ ? (v->binary ? "rb" : "r") : (v->binary ? "wb" : "w"));
    return v->f ? 0 : 65;
}
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char* fmt, ...) {
    if (!v->f) return 65;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(v->f, fmt, ap);
    va_end(ap);
    return 0;
}
PetscErrorCode PetscViewerDestroy(PetscViewer* v) {
    if (v && *v) {
        if ((*v)->f) fclose((*v)->f);
        delete *v;
        *v = nullptr;
    }
    return 0;
}
PetscErrorCode VecView(Vec v, PetscViewer viewer) {
    if (!viewer || !viewer->f) return 65;
    uint32_t hdr[2] = {bswap32(1211214u), bswap32((uint32_t)v->n)};
    fwrite(hdr, 4, 2, viewer->f);
    for (PetscInt i = 0; i < v->n; i++) {
        uint64_t u;
        memcpy(&u, &v->a[i], 8);
        u = bswap64(u);
        fwrite(&u, 8, 1, viewer->f);
    }
    return 0;
}
PetscErrorCode VecLoad(Vec v, PetscViewer viewer) {
    if (!viewer || !viewer->f) return 65;
    uint32_t hdr[2];
    if (fread(hdr, 4, 2, viewer->f) != 2) return 66;
    if (bswap32(hdr[0]) != 1211214u || (PetscInt)bswap32(hdr[1]) != v->n) return 67;
    for (PetscInt i = 0; i < v->n; i++) {
        uint64_t u;
        if (fread(&u, 8, 1, viewer->f) != 1) return 66;
        u = bswap64(u);
        memcpy(&v->a[i], &u, 8);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Mat
// ------------------------------------------------------------------------------------------------
static Mat mat_new(PetscInt m, PetscInt n) {
    auto* A = new _p_Mat();
    A->m = m;
    A->n = n;
    A->rows.resize((size_t)m);
    return A;
}
PetscErrorCode MatCreateAIJ(MPI_Comm, PetscInt m, PetscInt n, PetscInt, PetscInt, PetscInt, const PetscInt*, PetscInt, const PetscInt*,
                            Mat* A) {
    *A = mat_new(m, n);
    return 0;
}
PetscErrorCode MatSetLocalToGlobalMapping(Mat, ISLocalToGlobalMapping, ISLocalToGlobalMapping) { return 0; }
PetscErrorCode MatSetValuesLocal(Mat A, PetscInt m, const PetscInt* rows, PetscInt n, const PetscInt* cols, const PetscScalar* v,
                                 InsertMode mode) {
    for (PetscInt i = 0; i < m; i++) {
        if (rows[i] < 0) continue;  // PETSc ignores negative indices
        if (rows[i] >= A->m) return 63;
        A->csr_valid = false;
        auto& row = A->rows[(size_t)rows[i]];
        for (PetscInt j = 0; j < n; j++) {
            if (cols[j] < 0) continue;
            if (cols[j] >= A->n) return 63;
            if (mode == ADD_VALUES) row[cols[j]] += v[i * n + j];
            else row[cols[j]] = v[i * n + j];
        }
    }
    return 0;
}
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatZeroEntries(Mat A) {
    A->csr_valid = false;
    for (auto& r : A->rows)
        for (auto& kv : r) kv.second = 0.0;
    return 0;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) {
    if (x->n != A->n || y->n != A->m) return 60;
    std::vector<PetscScalar> out((size_t)A->m);
    A->apply(x->a, out.data());
    memcpy(y->a, out.data(), sizeof(PetscScalar) * (size_t)A->m);
    return 0;
}
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y) {
    if (x->n != A->m || y->n != A->n) return 60;
    std::vector<PetscScalar> out((size_t)A->n, 0.0);
    for (PetscInt i = 0; i < A->m; i++)
        for (const auto& kv : A->rows[(size_t)i]) out[(size_t)kv.first] += kv.second * x->a[i];
    memcpy(y->a, out.data(), sizeof(PetscScalar) * (size_t)A->n);
    return 0;
}
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r) {  // A <- diag(l) A diag(r)
    A->csr_valid = false;
    for (PetscInt i = 0; i < A->m; i++)
        for (auto& kv : A->rows[(size_t)i]) kv.second *= (l ? l->a[i] : 1.0) * (r ? r->a[kv.first] : 1.0);
    return 0;
}
PetscErrorCode MatDiagonalSet(Mat A, Vec d, InsertMode mode) {
    A->csr_valid = false;
    for (PetscInt i = 0; i < A->m && i < A->n; i++) {
        if (mode == ADD_VALUES) A->rows[(size_t)i][i] += d->a[i];
        else A->rows[(size_t)i][i] = d->a[i];
    }
    return 0;
}
PetscErrorCode MatDestroy(Mat* A) {
    if (A && *A) { delete *A; *A = nullptr; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// DMDA on one process: no ghost points (non-periodic), local numbering == natural numbering
// ------------------------------------------------------------------------------------------------
static DM_DA* DA(DM da) { return static_cast<DM_DA*>(da->data); }
PetscErrorCode DMDACreate3d(MPI_Comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz, DMDAStencilType st, PetscInt M,
                            PetscInt N, PetscInt P, PetscInt, PetscInt, PetscInt, PetscInt dof, PetscInt s, const PetscInt*,
                            const PetscInt*, const PetscInt*, DM* da) {
    if (bx != DM_BOUNDARY_NONE || by != DM_BOUNDARY_NONE || bz != DM_BOUNDARY_NONE) return 56;  // not needed by the reference
    auto* d = (DM_DA*)calloc(1, sizeof(DM_DA));
    d->M = M; d->N = N; d->P = P; d->dof = dof; d->s = s;
    d->bx = bx; d->by = by; d->bz = bz; d->stencil = st;
    d->elementtype = DMDA_ELEMENT_P1;  // PETSc's default; the reference switches to Q1
    d->lx[0] = M; d->ly[0] = N; d->lz[0] = P;
    auto* dm = new _p_DM();
    dm->data = d;
    *da = dm;
    return 0;
}
PetscErrorCode DMSetFromOptions(DM) { return 0; }
PetscErrorCode DMSetUp(DM) { return 0; }
PetscErrorCode DMDestroy(DM* da) {
    if (da && *da) {
        DM_DA* d = DA(*da);
        if (d) {
            if (d->coords) VecDestroy(&d->coords);
            free(d->e);
            free(d);
        }
        delete *da;
        *da = nullptr;
    }
    return 0;
}
PetscErrorCode DMDASetUniformCoordinates(DM da, PetscReal xmin, PetscReal xmax, PetscReal ymin, PetscReal ymax, PetscReal zmin,
                                         PetscReal zmax) {
    DM_DA* d = DA(da);
    if (d->coords) VecDestroy(&d->coords);
    VecCreateSeq(0, 3 * d->M * d->N * d->P, &d->coords);
    // non-periodic: h = (max - min) / (points - 1), coordinate = min + h * index
    const PetscReal hx = d->M > 1 ? (xmax - xmin) / (d->M - 1) : 0.0, hy = d->N > 1 ? (ymax - ymin) / (d->N - 1) : 0.0,
                    hz = d->P > 1 ? (zmax - zmin) / (d->P - 1) : 0.0;
    PetscInt c = 0;
    for (PetscInt k = 0; k < d->P; k++)
        for (PetscInt j = 0; j < d->N; j++)
            for (PetscInt i = 0; i < d->M; i++) {
                d->coords->a[c++] = xmin + hx * i;
                d->coords->a[c++] = ymin + hy * j;
                d->coords->a[c++] = zmin + hz * k;
            }
    return 0;
}
PetscErrorCode DMDASetElementType(DM da, DMDAElementType t) { DA(da)->elementtype = t; return 0; }
PetscErrorCode DMDAGetInfo(DM da, PetscInt* dim, PetscInt* M, PetscInt* N, PetscInt* P, PetscInt* m, PetscInt* n, PetscInt* p,
                           PetscInt* dof, PetscInt* s, DMBoundaryType* bx, DMBoundaryType* by, DMBoundaryType* bz, DMDAStencilType* st) {
    DM_DA* d = DA(da);
    if (dim) *dim = 3;
    if (M) *M = d->M;
    if (N) *N = d->N;
    if (P) *P = d->P;
    if (m) *m = 1;
    if (n) *n = 1;
    if (p) *p = 1;
    if (dof) *dof = d->dof;
    if (s) *s = d->s;
    if (bx) *bx = d->bx;
    if (by) *by = d->by;
    if (bz) *bz = d->bz;
    if (st) *st = d->stencil;
    return 0;
}
PetscErrorCode DMDAGetCorners(DM da, PetscInt* x, PetscInt* y, PetscInt* z, PetscInt* m, PetscInt* n, PetscInt* p) {
    DM_DA* d = DA(da);
    if (x) *x = 0;
    if (y) *y = 0;
    if (z) *z = 0;
    if (m) *m = d->M;
    if (n) *n = d->N;
    if (p) *p = d->P;
    return 0;
}
PetscErrorCode DMDAGetGhostCorners(DM da, PetscInt* x, PetscInt* y, PetscInt* z, PetscInt* m, PetscInt* n, PetscInt* p) {
    return DMDAGetCorners(da, x, y, z, m, n, p);
}
PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo* info) {
    DM_DA* d = DA(da);
    memset(info, 0, sizeof *info);
    info->dim = 3; info->dof = d->dof; info->sw = d->s;
    info->mx = d->M; info->my = d->N; info->mz = d->P;
    info->xm = info->gxm = d->M; info->ym = info->gym = d->N; info->zm = info->gzm = d->P;
    info->bx = d->bx; info->by = d->by; info->bz = d->bz; info->st = d->stencil;
    info->da = da;
    return 0;
}
PetscErrorCode DMDAGetOwnershipRanges(DM da, const PetscInt** lx, const PetscInt** ly, const PetscInt** lz) {
    DM_DA* d = DA(da);
    if (lx) *lx = d->lx;
    if (ly) *ly = d->ly;
    if (lz) *lz = d->lz;
    return 0;
}
PetscErrorCode DMDAGetElements(DM, PetscInt*, PetscInt*, const PetscInt**) { return 56; }  // the reference brings its own
PetscErrorCode DMDARestoreElements(DM, PetscInt*, PetscInt*, const PetscInt**) { return 0; }
PetscErrorCode DMGetCoordinatesLocal(DM da, Vec* c) { *c = DA(da)->coords; return *c ? 0 : 73; }
PetscErrorCode DMCreateGlobalVector(DM da, Vec* v) { DM_DA* d = DA(da); return VecCreateSeq(0, d->dof * d->M * d->N * d->P, v); }
PetscErrorCode DMCreateLocalVector(DM da, Vec* v) { return DMCreateGlobalVector(da, v); }
PetscErrorCode DMGlobalToLocalBegin(DM, Vec g, InsertMode mode, Vec l) {
    if (mode == ADD_VALUES) return VecAXPY(l, 1.0, g);
    return VecCopy(g, l);
}
PetscErrorCode DMGlobalToLocalEnd(DM, Vec, InsertMode, Vec) { return 0; }
PetscErrorCode DMCreateMatrix(DM da, Mat* A) {
    DM_DA* d = DA(da);
    const PetscInt n = d->dof * d->M * d->N * d->P;
    *A = mat_new(n, n);
    return 0;
}
PetscErrorCode DMGetLocalToGlobalMapping(DM, ISLocalToGlobalMapping* m) { *m = &g_l2g; return 0; }
PetscErrorCode DMCoarsenHierarchy(DM da, PetscInt nlevels, DM* dac) {
    DM cur = da;
    for (PetscInt l = 0; l < nlevels; l++) {
        DM_DA* d = DA(cur);
        if (((d->M - 1) % 2) || ((d->N - 1) % 2) || ((d->P - 1) % 2)) return 60;
        DMDACreate3d(0, d->bx, d->by, d->bz, d->stencil, (d->M - 1) / 2 + 1, (d->N - 1) / 2 + 1, (d->P - 1) / 2 + 1, 1, 1, 1, d->dof, d->s,
                     nullptr, nullptr, nullptr, &dac[l]);
        cur = dac[l];
    }
    return 0;
}
PetscErrorCode DMCreateInterpolation(DM, DM, Mat* R, Vec* scale) {
    *R = mat_new(0, 0);  // the multigrid preconditioner lives inside PETSc; KSPSolve below does not use it
    if (scale) *scale = nullptr;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// KSP / PC: Jacobi-PCG to 1e-13 ||b|| whatever solver the caller configured (see petsc.h)
// ------------------------------------------------------------------------------------------------
PetscErrorCode KSPCreate(MPI_Comm, KSP* ksp) {
    *ksp = new _p_KSP();
    (*ksp)->pc = new _p_PC();
    return 0;
}
PetscErrorCode KSPDestroy(KSP* ksp) {
    if (ksp && *ksp) {
        if ((*ksp)->pc) {
            for (KSP s : (*ksp)->pc->sub) KSPDestroy(&s);
            delete (*ksp)->pc;
        }
        delete *ksp;
        *ksp = nullptr;
    }
    return 0;
}
PetscErrorCode KSPSetType(KSP ksp, KSPType t) { ksp->type = t; return 0; }
PetscErrorCode KSPGetType(KSP ksp, KSPType* t) { *t = ksp->type.c_str(); return 0; }
PetscErrorCode KSPGMRESSetRestart(KSP ksp, PetscInt r) { ksp->restart = r; return 0; }
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal atol, PetscReal dtol, PetscInt maxits) {
    if (rtol != (PetscReal)PETSC_DEFAULT) ksp->rtol = rtol;
    if (atol != (PetscReal)PETSC_DEFAULT) ksp->atol = atol;
    if (dtol != (PetscReal)PETSC_DEFAULT) ksp->dtol = dtol;
    if (maxits != PETSC_DEFAULT) ksp->maxits = maxits;
    return 0;
}
PetscErrorCode KSPGetTolerances(KSP ksp, PetscReal* rtol, PetscReal* atol, PetscReal* dtol, PetscInt* maxits) {
    if (rtol) *rtol = ksp->rtol;
    if (atol) *atol = ksp->atol;
    if (dtol) *dtol = ksp->dtol;
    if (maxits) *maxits = ksp->maxits;
    return 0;
}
PetscErrorCode KSPSetInitialGuessNonzero(KSP ksp, PetscBool flg) { ksp->guess_nonzero = flg; return 0; }
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat) { ksp->A = A; return 0; }
PetscErrorCode KSPSetFromOptions(KSP) { return 0; }
PetscErrorCode KSPSetUp(KSP) { return 0; }
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt* its) { *its = ksp->its; return 0; }
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal* rnorm) { *rnorm = ksp->rnorm; return 0; }
PetscErrorCode KSPGetPC(KSP ksp, PC* pc) { *pc = ksp->pc; return 0; }
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x) {
    Mat A = ksp->A;
    if (!A || A->m != A->n || b->n != A->m || x->n != A->m) return 60;
    const PetscInt n = A->m;
    if (!ksp->guess_nonzero) VecSet(x, 0.0);
    std::vector<PetscScalar> r((size_t)n), z((size_t)n), p((size_t)n), q((size_t)n), dinv((size_t)n);
    for (PetscInt i = 0; i < n; i++) {
        auto it = A->rows[(size_t)i].find(i);
        const PetscScalar d = it == A->rows[(size_t)i].end() ? 0.0 : it->second;
        dinv[(size_t)i] = d != 0.0 ? 1.0 / d : 1.0;
    }
    auto apply = [&](const PetscScalar* u, PetscScalar* y) { A->apply(u, y); };
    PetscReal bnorm = 0;
    for (PetscInt i = 0; i < n; i++) bnorm += b->a[i] * b->a[i];
    bnorm = sqrt(bnorm);
    apply(x->a, q.data());
    PetscReal rz = 0, rr = 0;
    // (the loops below run under OpenMP when the build enables it: long fixture runs, tests/golden/make_trajectory_golden.py)
#pragma omp parallel for schedule(static) reduction(+ : rz, rr)
    for (PetscInt i = 0; i < n; i++) {
        r[(size_t)i] = b->a[i] - q[(size_t)i];
        z[(size_t)i] = dinv[(size_t)i] * r[(size_t)i];
        p[(size_t)i] = z[(size_t)i];
        rz += r[(size_t)i] * z[(size_t)i];
        rr += r[(size_t)i] * r[(size_t)i];
    }
    const PetscReal tol = 1e-13 * bnorm;
    PetscInt it = 0;
    while (sqrt(rr) > tol && it < 200000) {
        apply(p.data(), q.data());
        PetscReal pq = 0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
        for (PetscInt i = 0; i < n; i++) pq += p[(size_t)i] * q[(size_t)i];
        if (!(pq > 0)) break;
        const PetscReal alpha = rz / pq;
        PetscReal rz_new = 0;
        rr = 0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
        for (PetscInt i = 0; i < n; i++) {
            x->a[i] += alpha * p[(size_t)i];
            r[(size_t)i] -= alpha * q[(size_t)i];
            z[(size_t)i] = dinv[(size_t)i] * r[(size_t)i];
            rz_new += r[(size_t)i] * z[(size_t)i];
            rr += r[(size_t)i] * r[(size_t)i];
        }
        const PetscReal beta = rz_new / rz;
        rz = rz_new;
#pragma omp parallel for schedule(static)
        for (PetscInt i = 0; i < n; i++) p[(size_t)i] = z[(size_t)i] + beta * p[(size_t)i];
        it++;
    }
    ksp->its = it;
    ksp->rnorm = sqrt(rr);
    return 0;
}
PetscErrorCode PCSetType(PC pc, PCType t) { pc->type = t; return 0; }
PetscErrorCode PCGetType(PC pc, PCType* t) { *t = pc->type.c_str(); return 0; }
PetscErrorCode PCSetReusePreconditioner(PC, PetscBool) { return 0; }
PetscErrorCode PCMGSetLevels(PC pc, PetscInt levels, MPI_Comm*) {
    for (KSP s : pc->sub) KSPDestroy(&s);
    pc->sub.clear();
    for (PetscInt l = 0; l < levels; l++) {
        KSP k;
        KSPCreate(0, &k);
        pc->sub.push_back(k);
    }
    return 0;
}
PetscErrorCode PCMGSetType(PC, PCMGType) { return 0; }
PetscErrorCode PCMGSetCycleType(PC, PCMGCycleType) { return 0; }
PetscErrorCode PCMGSetGalerkin(PC, PCMGGalerkinType) { return 0; }
PetscErrorCode PCMGSetInterpolation(PC, PetscInt, Mat) { return 0; }
PetscErrorCode PCMGGetSmoother(PC pc, PetscInt l, KSP* ksp) {
    if (l < 0 || l >= (PetscInt)pc->sub.size()) return 63;
    *ksp = pc->sub[(size_t)l];
    return 0;
}
PetscErrorCode PCMGGetCoarseSolve(PC pc, KSP* ksp) { return PCMGGetSmoother(pc, 0, ksp); }

// ------------------------------------------------------------------------------------------------
// MPI on one process
// ------------------------------------------------------------------------------------------------
struct plite_datatype {
    int size;                         // bytes of one element (basic types) / of the old type (vector types)
    int count, blocklength, stride;   // vector types: count blocks of blocklength elements, stride elements apart
    bool vector;
};
static plite_datatype t_double{8, 0, 0, 0, false}, t_int{4, 0, 0, 0, false}, t_ulong{8, 0, 0, 0, false}, t_float{4, 0, 0, 0, false},
    t_char{1, 0, 0, 0, false};
MPI_Datatype MPI_DOUBLE = &t_double, MPI_INT = &t_int, MPI_UNSIGNED_LONG = &t_ulong, MPI_FLOAT = &t_float, MPI_CHAR = &t_char;
struct plite_file {
    FILE* f;
    MPI_Offset disp;
    MPI_Datatype filetype;
    MPI_Offset pos;  // bytes already written through the current view (counted in visible bytes)
};
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return 0; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return 0; }
int MPI_Allreduce(const void* s, void* r, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
    if (s != r) memmove(r, s, (size_t)count * (size_t)type->size);
    return 0;
}
int MPI_Allgather(const void* s, int scount, MPI_Datatype stype, void* r, int, MPI_Datatype, MPI_Comm) {
    if (s != r) memmove(r, s, (size_t)scount * (size_t)stype->size);
    return 0;
}
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return 0; }
int MPI_Barrier(MPI_Comm) { return 0; }
int MPI_Abort(MPI_Comm, int code) { exit(code ? code : 1); }
double MPI_Wtime(void) {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int MPI_Type_size(MPI_Datatype type, int* size) { *size = type->size; return 0; }
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype) {
    *newtype = new plite_datatype{oldtype->size, count, blocklength, stride, true};
    return 0;
}
int MPI_Type_commit(MPI_Datatype*) { return 0; }
int MPI_Type_free(MPI_Datatype* type) {
    if (type && *type && (*type)->vector) { delete *type; *type = nullptr; }
    return 0;
}
int MPI_File_open(MPI_Comm, const char* filename, int amode, MPI_Info, MPI_File* fh) {
    FILE* f = fopen(filename, "r+b");
    if (!f && (amode & MPI_MODE_CREATE)) f = fopen(filename, "w+b");
    if (!f) return 1;
    *fh = new plite_file{f, 0, MPI_CHAR, 0};
    return 0;
}
int MPI_File_close(MPI_File* fh) {
    if (fh && *fh) { fclose((*fh)->f); delete *fh; *fh = nullptr; }
    return 0;
}
int MPI_File_delete(const char* filename, MPI_Info) { return unlink(filename) == 0 ? 0 : 1; }
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype, MPI_Datatype filetype, const char*, MPI_Info) {
    fh->disp = disp;
    fh->filetype = filetype;
    fh->pos = 0;
    return 0;
}
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status*) {
    const char* src = static_cast<const char*>(buf);
    MPI_Offset nbytes = (MPI_Offset)count * type->size;
    if (!fh->filetype->vector) {
        fseeko(fh->f, (off_t)(fh->disp + fh->pos), SEEK_SET);
        if (fwrite(src, 1, (size_t)nbytes, fh->f) != (size_t)nbytes) return 1;
        fh->pos += nbytes;
        return 0;
    }
    // strided view: the visible bytes are blocks of blocklength elements every stride elements
    const MPI_Offset blk = (MPI_Offset)fh->filetype->blocklength * fh->filetype->size;
    const MPI_Offset str = (MPI_Offset)fh->filetype->stride * fh->filetype->size;
    while (nbytes > 0) {
        const MPI_Offset b = fh->pos / blk, in = fh->pos % blk;
        const MPI_Offset chunk = (blk - in) < nbytes ? (blk - in) : nbytes;
        fseeko(fh->f, (off_t)(fh->disp + b * str + in), SEEK_SET);
        if (fwrite(src, 1, (size_t)chunk, fh->f) != (size_t)chunk) return 1;
        src += chunk; nbytes -= chunk; fh->pos += chunk;
    }
    return 0;
}
int MPI_File_write_all(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* st) {
    return MPI_File_write(fh, buf, count, type, st);
}
