// comm.cu -- z-slab halo exchange and scalar all-reduce over NCCL (NVLink 5 / NVSwitch).
// Replaces DMGlobalToLocalBegin/End (LinearElasticity.cc:249-250,317-318,389-390), the VecScatter
// inside MatMult, and the MPI_Allreduce calls of LinearElasticity.cc:283,429 / Filter.cc:221.
// One nodal z-plane (all three components) is contiguous in the stored layout, so each
// neighbour exchange is a single ncclSend/ncclRecv pair of pstride doubles.
#include "common.cuh"

namespace tob {

int halo_nodes(tob_ctx_impl* c, const LevelGrid& g, double* v) {
    if (c->desc.nranks == 1) return 0;
    const int r = c->desc.rank, n = c->desc.nranks;
    const size_t cnt = (size_t)g.pstride;
    TOB_NCCL(ncclGroupStart());
    if (r > 0) {  // lower neighbour: send my first owned plane, receive its last owned plane into local plane 0
        TOB_NCCL(ncclSend(v + (long long)g.zl_own0() * g.pstride, cnt, ncclDouble, r - 1, c->comm, c->stream));
        TOB_NCCL(ncclRecv(v, cnt, ncclDouble, r - 1, c->comm, c->stream));
    }
    if (r < n - 1) {  // upper neighbour
        TOB_NCCL(ncclSend(v + (long long)(g.zl_own1() - 1) * g.pstride, cnt, ncclDouble, r + 1, c->comm, c->stream));
        TOB_NCCL(ncclRecv(v + (long long)(g.nzl() - 1) * g.pstride, cnt, ncclDouble, r + 1, c->comm, c->stream));
    }
    TOB_NCCL(ncclGroupEnd());
    return 0;
}

// Element field with local layers: owned [own0,own1); recv_lo layers below own0 and recv_hi above
// own1 are filled from the z-neighbours; send_down / send_up are the numbers of my boundary
// layers the lower / upper neighbour stores as its ghosts (known analytically by the caller).
int halo_elems(tob_ctx_impl* c, double* v, long long layer_stride, int own0, int own1, int recv_lo, int recv_hi,
               int send_down, int send_up) {
    if (c->desc.nranks == 1) return 0;
    const int r = c->desc.rank, n = c->desc.nranks;
    if (send_down > own1 - own0 || send_up > own1 - own0) TOB_FAIL("ghost width exceeds the slab thickness");
    TOB_NCCL(ncclGroupStart());
    if (r > 0) {
        if (send_down) TOB_NCCL(ncclSend(v + (long long)own0 * layer_stride, (size_t)send_down * layer_stride, ncclDouble, r - 1, c->comm, c->stream));
        if (recv_lo) TOB_NCCL(ncclRecv(v + (long long)(own0 - recv_lo) * layer_stride, (size_t)recv_lo * layer_stride, ncclDouble, r - 1, c->comm, c->stream));
    }
    if (r < n - 1) {
        if (send_up) TOB_NCCL(ncclSend(v + (long long)(own1 - send_up) * layer_stride, (size_t)send_up * layer_stride, ncclDouble, r + 1, c->comm, c->stream));
        if (recv_hi) TOB_NCCL(ncclRecv(v + (long long)own1 * layer_stride, (size_t)recv_hi * layer_stride, ncclDouble, r + 1, c->comm, c->stream));
    }
    TOB_NCCL(ncclGroupEnd());
    return 0;
}

int allreduce_sum(tob_ctx_impl* c, double* d_vals, int n) {
    if (c->desc.nranks == 1) return 0;
    TOB_NCCL(ncclAllReduce(d_vals, d_vals, n, ncclDouble, ncclSum, c->comm, c->stream));
    return 0;
}
int allreduce_max(tob_ctx_impl* c, double* d_vals, int n) {
    if (c->desc.nranks == 1) return 0;
    TOB_NCCL(ncclAllReduce(d_vals, d_vals, n, ncclDouble, ncclMax, c->comm, c->stream));
    return 0;
}

}  // namespace tob
