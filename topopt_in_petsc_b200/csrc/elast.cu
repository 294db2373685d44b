// elast.cu -- state problem: BC/load, matrix-free PCG, objective and sensitivities.
// Replaces the inside of class LinearElasticity (LinearElasticity.cc:12-549): the assembled AIJ
// matrix + KSPFGMRES/PCMG become a matrix-free hex8 operator + PCG (Jacobi or GMG V-cycle).
#include "elast.cuh"
#include "vecops.cuh"

namespace tob {

// scalar slots
enum { S_ZERO = 0, S_ONE = 1, S_PQ = 2, S_BB = 3, S_PAIR_A = 4, S_PAIR_B = 6, S_OBJ = 8, S_TMP = 10 };

static int init_const_slots(tob_ctx_impl* c) {
    const double v[2] = {0.0, 1.0};
    TOB_CUDA(cudaMemcpyAsync(c->d_scalars + S_ZERO, v, sizeof v, cudaMemcpyHostToDevice, c->stream));
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int elast_apply(tob_elast_impl* h, const double* u, double* y, int epi, bool mask_in, int kernel, const double* a0,
                int* nparts) {
    SpmvArgs a;
    a.kernel = kernel ? kernel : h->opts.spmv_kernel;
    const double* u_orig = nullptr;
    if (mask_in && a.kernel != SPMV_NAIVE) {
        u_orig = u;
        // public SpMV on an arbitrary vector: apply N first (inside the solver fixed dofs are exactly 0)
        TOB_CHECK(vec_pointwise_mult(h->c, h->Z, u, h->NFREE, h->g->nvec));
        u = h->Z;
        mask_in = false;
    }
    a.u = u; a.y = y; a.E = h->E_st; a.mask = h->mask; a.epi = epi; a.mask_in = mask_in;
    a.a0 = a0;
    // Jacobi-PCG on one rank: working set >> L2, see common.cuh (the smaller per-rank sets of N > 1 mostly fit)
    if (epi == EPI_DOT && h->opts.pc == 0 && h->c->desc.nranks == 1) a.l2hint = h->c->l2hint;
    if (epi == EPI_CGDOT) {
        a.cg = h->sr_cg; a.cgpar = h->sr_par; a.upart = h->sr_upart; a.nbU = h->sr_nbU; a.pe = &h->sr_pe;
        if (h->c->desc.nranks == 1) a.l2hint = h->c->l2hint & 1;  // modulus tiles evict_first
    }
    h->stats.spmv_calls++;
    constexpr int kMaxPairs = 2048;
    // profiling mode 1 times every launch, mode >1 every mode-th launch (event pairs around every launch cost ~6 %
    // of a Jacobi-PCG iteration; sampling keeps the live roofline measurement out of the way of the timed step)
    const bool timed = h->profile && h->prof_used < kMaxPairs && (h->prof_stride <= 1 || (h->stats.spmv_calls % h->prof_stride) == 0);
    if (timed) {
        if ((int)h->prof_ev.size() < 2 * (h->prof_used + 1)) {
            cudaEvent_t e0, e1;
            TOB_CUDA(cudaEventCreate(&e0));
            TOB_CUDA(cudaEventCreate(&e1));
            h->prof_ev.push_back(e0);
            h->prof_ev.push_back(e1);
        }
        TOB_CUDA(cudaEventRecord(h->prof_ev[2 * h->prof_used], h->c->stream));
    }
    int rc = hex8_spmv(h->c, *h->g, h->op, a, nparts);
    // pre-masked input: the kernel returned 0 on clamped dofs, the operator wants (I-N) u there
    if (rc == 0 && u_orig) rc = vec_restore_fixed(h->c, y + h->g->own_off, u_orig + h->g->own_off, h->NFREE + h->g->own_off, h->g->own_len);
    if (timed) {
        TOB_CUDA(cudaEventRecord(h->prof_ev[2 * h->prof_used + 1], h->c->stream));
        h->prof_used++;
    }
    return rc;
}

int elast_profile_collect(tob_elast_impl* h) {
    // called after the solve's final synchronisation: sum the recorded pairs
    for (int i = 0; i < h->prof_used; i++) {
        float ms = 0;
        TOB_CUDA(cudaEventSynchronize(h->prof_ev[2 * i + 1]));
        TOB_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
        h->stats.spmv_seconds += ms * 1e-3;
        h->stats.spmv_timed++;
    }
    h->prof_used = 0;
    return 0;
}

int elast_set_density(tob_elast_impl* h, const double* xphys_nat_dev, double Emin, double Emax, double penal) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    TOB_CHECK(nat_to_stored_elem(c, g, xphys_nat_dev, h->xphys_st));
    // ghost element layer above (owned by the upper neighbour): it sends its first owned layer down
    const int own0 = g.ke0 - g.ge0, own1 = g.ke1 - g.ge0;
    const int recv_hi = g.nzle() - own1;
    const int send_down = c->desc.rank > 0 ? 1 : 0;
    TOB_CHECK(halo_elems(c, h->xphys_st, g.estride, own0, own1, 0, recv_hi, send_down, 0));
    TOB_CHECK(simp_modulus(c, h->xphys_st, h->E_st, g.evec, Emin, Emax, penal));
    TOB_CHECK(hex8_diag_inv(c, g, h->op, h->E_st, h->mask, h->DINV, true));
    TOB_CHECK(vec_d2f(c, h->DINV, h->DINV32, g.nvec));
    h->Emin = Emin; h->Emax = Emax; h->penal = penal;
    h->have_density = true;
    h->E32_valid = false;
    if (h->mg) TOB_CHECK(gmg_update(h));
    return 0;
}


// ---- single-reduction Jacobi-PCG (cgsr.cuh): 2 launches per iteration, one reduction, halo fused into the producer ------
static int sr_setup(tob_elast_impl* h) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    if (h->sr_cg) return 0;
    TOB_CHECK(peer_vec_alloc(c, sizeof(double) * g.nvec, &h->sr_u));
    TOB_CUDA(cudaMalloc(&h->sr_cg, sizeof(CgDev)));
    TOB_CUDA(cudaMemsetAsync(h->sr_cg, 0, sizeof(CgDev), c->stream));
    TOB_CUDA(cudaMalloc(&h->sr_upart, sizeof(double) * 2 * (size_t)c->num_sms * 8));
    CgPeers& pe = h->sr_pe;
    pe = CgPeers();
    pe.n = c->desc.nranks; pe.rank = c->desc.rank;
    if (getenv("TOB_CG_TRACE") && atoi(getenv("TOB_CG_TRACE"))) {  // diagnostic: device time stamps of the U/S pairs (cgsr.cuh)
        pe.trace_cap = 4096;
        TOB_CUDA(cudaMalloc(&pe.trace, sizeof(unsigned long long) * 8 * pe.trace_cap));
        TOB_CUDA(cudaMemsetAsync(pe.trace, 0, sizeof(unsigned long long) * 8 * pe.trace_cap, c->stream));
    }
    if (pe.n > 1) {
        pe.arena = c->peer.arena;
        for (int k = 0; k < pe.n; k++) pe.peer[k] = c->peer.peer[k];
        if (pe.rank > 0) {  // the lower neighbour's TOP ghost plane = its last stored plane
            int kn0 = 0, kn1 = 0;
            TOB_CHECK(tob_slab_partition(g.nzn, pe.n, pe.rank - 1, &kn0, &kn1, nullptr, nullptr));
            const int gn0 = kn0 > 0 ? kn0 - 1 : 0, gn1 = kn1 + 1;
            pe.u_lo_ghost = h->sr_u.peer[pe.rank - 1] + (long long)(gn1 - gn0 - 1) * g.pstride;
        }
        if (pe.rank < pe.n - 1) pe.u_hi_ghost = h->sr_u.peer[pe.rank + 1];  // its plane 0 is its bottom ghost plane
        // nobody may store into a peer's buffer before that peer has zeroed it
        double* d_dummy = c->d_scalars + kScalarSlots - 1;
        TOB_NCCL(ncclAllReduce(d_dummy, d_dummy, 1, ncclDouble, ncclSum, c->comm, c->stream));
    }
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static void sr_destroy(tob_elast_impl* h) {
    if (!h->sr_cg) return;
    if (h->sr_graph) { cudaGraphExecDestroy(h->sr_graph); h->sr_graph = nullptr; }
    peer_vec_free(h->c, &h->sr_u);
    cudaFree(h->sr_cg); h->sr_cg = nullptr;
    cudaFree(h->sr_upart); h->sr_upart = nullptr;
    if (h->sr_pe.trace) { cudaFree(h->sr_pe.trace); h->sr_pe.trace = nullptr; }
}

// R holds b - A U on entry; iterates until ||r|| <= tol.  *it / *rnorm as in elast_pcg.
static int elast_pcg_sr(tob_elast_impl* h, double tol, double bnorm, int* it_out, double* rnorm_out) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    const tob_solver_opts& o = h->opts;
    TOB_CHECK(sr_setup(h));
    double* u = h->sr_u.mine;
    CgState st0{1.0, 1.0, 1.0, c->cg_epoch};
    TOB_CUDA(cudaMemcpyAsync(&h->sr_cg->st[0], &st0, sizeof st0, cudaMemcpyHostToDevice, c->stream));
    TOB_CHECK(vec_zero(c, h->P, g.nvec));
    TOB_CHECK(vec_zero(c, h->Z, g.nvec));
    int np = 0;
    // set-up pair: u = D^-1 r (+ halo), w = A u, sums
    TOB_CHECK(cg_update(c, g, h->sr_cg, 0, 0, h->U, h->R, u, h->Q, h->P, h->Z, h->DINV32, h->sr_upart, &h->sr_nbU, h->sr_pe));
    c->cg_epoch++;
    h->sr_par = 1;
    TOB_CHECK(elast_apply(h, u, h->Q, EPI_CGDOT, false, 0, nullptr, &np));
    int par = 1, it = 0;
    static const bool lazyx = !(getenv("TOB_PCG_LAZYX") && atoi(getenv("TOB_PCG_LAZYX")) == 0);
    auto one_iteration = [&](int xm) -> int {  // xm: 1 = even, 2 = odd iteration of a pair (lazy x update, cgsr.cu)
        TOB_CHECK(cg_update(c, g, h->sr_cg, par, 1, h->U, h->R, u, h->Q, h->P, h->Z, h->DINV32, h->sr_upart, &h->sr_nbU, h->sr_pe, lazyx ? xm : 0));
        c->cg_epoch++;
        par ^= 1;
        h->sr_par = par;
        TOB_CHECK(elast_apply(h, u, h->Q, EPI_CGDOT, false, 0, nullptr, &np));
        return 0;
    };
    const int check = o.check_every > 0 ? o.check_every : 10;
    const int blk = std::max(2, check & ~1);
    const bool graphed = !getenv("TOB_PCG_NOGRAPH") && check >= 4;
    const int stride_saved = h->prof_stride;
    if (graphed) h->prof_stride = 1;  // the two eager iterations of every round are the timed sample
    double rnorm = *rnorm_out;
    double* hs = c->h_scalars + 40;
    while (it < o.maxits) {
        TOB_CHECK(one_iteration(1));
        TOB_CHECK(one_iteration(2));
        it += 2;
        if (graphed) {
            if (!h->sr_graph || h->sr_graph_its != blk) {
                if (h->sr_graph) { cudaGraphExecDestroy(h->sr_graph); h->sr_graph = nullptr; }
                const bool prof = h->profile;
                h->profile = false;
                const long long l0 = c->launches;
                const unsigned long long e0 = c->cg_epoch;
                TOB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                int rc = 0;
                for (int k = 0; k < blk && !rc; k++) rc = one_iteration(1 + (k & 1));
                cudaGraph_t graph = nullptr;
                cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
                h->profile = prof;
                c->cg_epoch = e0;  // nothing ran during the capture
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) TOB_FAIL(std::string("PCG graph capture failed: ") + cudaGetErrorString(e));
                TOB_CUDA(cudaGraphInstantiate(&h->sr_graph, graph, 0));
                cudaGraphDestroy(graph);
                h->sr_graph_its = blk;
                h->sr_graph_launches = c->launches - l0;
                c->launches = l0;
            }
            TOB_CUDA(cudaGraphLaunch(h->sr_graph, c->stream));
            c->launches += h->sr_graph_launches;
            c->cg_epoch += blk;
            h->stats.spmv_calls += blk;
            it += blk;
        }
        // ||r||^2 as the last U saw it (the residual before its own update: one iteration behind, on the safe side)
        TOB_CUDA(cudaMemcpyAsync(hs, h->sr_cg->out, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TOB_CUDA(cudaStreamSynchronize(c->stream));
        rnorm = sqrt(hs[1]);
        if (!(rnorm == rnorm)) TOB_FAIL("PCG produced NaN (bad density field)");
        if (rnorm <= tol || rnorm > o.dtol * bnorm) break;
    }
    h->prof_stride = stride_saved;
    *it_out = it;
    *rnorm_out = rnorm;
    if (h->sr_pe.trace) {  // TOB_CG_TRACE: mean phase durations over the epochs still in the ring
        const int cap = h->sr_pe.trace_cap;
        std::vector<unsigned long long> t((size_t)cap * 8);
        TOB_CUDA(cudaMemcpy(t.data(), h->sr_pe.trace, t.size() * 8, cudaMemcpyDeviceToHost));
        double acc[6] = {0, 0, 0, 0, 0, 0};
        long long cnt = 0;
        const unsigned long long e1 = c->cg_epoch;
        for (unsigned long long e = (e1 > (unsigned long long)cap ? e1 - cap + 2 : 2); e + 1 < e1; e++) {
            const unsigned long long* a = &t[(e % cap) * 8];
            const unsigned long long* b = &t[((e + 1) % cap) * 8];
            if (!a[0] || !a[5] || !b[0] || a[5] < a[0] || b[0] < a[5]) continue;
            acc[0] += (double)(a[1] - a[0]);                     // U: wait for the sums
            acc[1] += (double)((a[2] ? a[2] : a[3]) - a[1]);     // U: body (to the last block / next entry on one rank)
            acc[2] += (double)(a[3] - (a[2] ? a[2] : a[3]));     // gap U -> S
            acc[3] += (double)(a[4] ? a[4] - a[3] : 0);          // S: wait for the ghost planes
            acc[4] += (double)(a[5] - (a[4] ? a[4] : a[3]));     // S: body + reduction
            acc[5] += (double)(b[0] - a[5]);                     // gap S -> next U
            cnt++;
        }
        if (cnt)
            fprintf(stderr, "[tob cg trace] rank %d: %lld iterations  U-wait %.2f  U-body %.2f  gap %.2f  S-wait %.2f  S-body %.2f  gap %.2f  (us, mean)\n",
                    c->desc.rank, cnt, acc[0] / cnt * 1e-3, acc[1] / cnt * 1e-3, acc[2] / cnt * 1e-3, acc[3] / cnt * 1e-3, acc[4] / cnt * 1e-3, acc[5] / cnt * 1e-3);
    }
    return 0;
}

// PCG on A U = B, warm-started from the resident U (KSPSetInitialGuessNonzero, LinearElasticity.cc:647)
int elast_pcg(tob_elast_impl* h) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    const tob_solver_opts& o = h->opts;
    const bool jac = (o.pc == 0) || !h->mg;
    TOB_CUDA(cudaEventRecord(c->ev0, c->stream));
    h->stats.iterations = 0;
    int np = 0;
    // r = b - A u   (fixed dofs of U are exactly zero: U was masked when set)
    TOB_CHECK(halo_nodes(c, g, h->U));
    TOB_CHECK(elast_apply(h, h->U, h->R, EPI_RESID, false, 0, h->B, &np));
    TOB_CHECK(finalize_sum(c, np, 1, S_TMP, true));
    double rr = 0, tmp[2];
    TOB_CHECK(read_scalars(c, S_TMP, 1, &rr));
    const double bnorm = h->bnorm;
    const double tol = fmax(o.rtol * bnorm, o.atol);  // KSPConvergedDefault: ||r|| <= max(rtol ||b||, atol)
    double rnorm = sqrt(rr);
    int it = 0;
    const bool fused = jac && c->desc.nranks == 1 && !getenv("TOB_PCG_UNFUSED");
    // Single-reduction variant: the default on several ranks (latency: 2 launches / 2 NVLink hops per iteration instead of
    // 6 / 3).  Its recurrences drift from the true residual below ~1e-9, so tight solves keep the classical iteration.
    // TOB_PCG_SR = 0 / 1 forces it off / on (1 also on one rank).
    const char* sr_env = getenv("TOB_PCG_SR");
    const bool sr_ok = jac && (o.spmv_kernel == 0 || o.spmv_kernel == SPMV_ZM16U4) && (c->desc.nranks == 1 || c->peer.enabled);
    const bool sr = sr_ok && (sr_env ? atoi(sr_env) != 0 : (c->desc.nranks > 1 && o.rtol >= 1e-8));
    if (rnorm > tol && bnorm > 0 && sr) {
        TOB_CHECK(elast_pcg_sr(h, tol, bnorm, &it, &rnorm));
    } else if (rnorm > tol && bnorm > 0 && fused) {
        // single-GPU Jacobi-PCG, 3 launches per iteration: the consumers reduce the producers' per-CTA partials
        // themselves (vecops.cu), no finalize kernels, alpha/beta never leave the device
        double* blk_cur = c->d_partials + (256 << 10);
        double* blk_nxt = c->d_partials + (512 << 10);
        int nbx = 0;
        TOB_CHECK(pcg_update_xr_f(c, g, h->U, h->R, h->R, h->R, h->DINV32, nullptr, 0, nullptr, blk_cur, &nbx));
        TOB_CHECK(pcg_update_p_f(c, g, h->P, h->R, h->DINV32, nullptr, nullptr, nullptr, nbx));
        const int check = o.check_every > 0 ? o.check_every : 10;
        const bool graphed1 = !getenv("TOB_PCG_NOGRAPH") && check >= 4;
        // iterations run in (even, odd) pairs with the lazy x update of vecops.cu (xm = 1, 2); TOB_PCG_LAZYX=0 switches it off
        static const bool lazyx = !(getenv("TOB_PCG_LAZYX") && atoi(getenv("TOB_PCG_LAZYX")) == 0);
        double* ab = c->d_scalars + 12;
        auto one_iteration1 = [&](double* bc, double* bn, int xm) -> int {
            TOB_CHECK(elast_apply(h, h->P, h->Q, EPI_DOT, false, 0, nullptr, &np));
            TOB_CHECK(pcg_update_xr_f(c, g, h->U, h->R, h->P, h->Q, h->DINV32, c->d_partials, np, bc, bn, &nbx, lazyx ? xm : 0, ab));
            TOB_CHECK(pcg_update_p_f(c, g, h->P, h->R, h->DINV32, bn, bc, ab, nbx));
            return 0;
        };
        const int stride_saved1 = h->prof_stride;
        if (graphed1) h->prof_stride = 1;  // the two eager iterations per round are the timed sample
        while (graphed1 && it < o.maxits) {
            // same scheme as the multi-GPU path below: 2 eager (timed) iterations, then a replayed graph of `check`
            TOB_CHECK(one_iteration1(blk_cur, blk_nxt, 1));
            TOB_CHECK(one_iteration1(blk_nxt, blk_cur, 2));
            it += 2;
            const int blk = check & ~1;
            if (!h->pcg_graph || h->pcg_graph_its != blk) {
                if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }
                const bool prof = h->profile;
                h->profile = false;
                const long long l0 = c->launches;
                TOB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                int rc = 0;
                for (int k = 0; k < blk && !rc; k += 2) {
                    rc = one_iteration1(blk_cur, blk_nxt, 1);
                    if (!rc) rc = one_iteration1(blk_nxt, blk_cur, 2);
                }
                cudaGraph_t graph = nullptr;
                cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
                h->profile = prof;
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) TOB_FAIL(std::string("PCG graph capture failed: ") + cudaGetErrorString(e));
                TOB_CUDA(cudaGraphInstantiate(&h->pcg_graph, graph, 0));
                cudaGraphDestroy(graph);
                h->pcg_graph_its = blk;
                h->pcg_graph_launches = c->launches - l0;
                c->launches = l0;
            }
            TOB_CUDA(cudaGraphLaunch(h->pcg_graph, c->stream));
            c->launches += h->pcg_graph_launches;
            it += blk;
            TOB_CHECK(finalize_parts(c, blk_cur, nbx, S_TMP));
            TOB_CHECK(read_scalars(c, S_TMP, 2, tmp));
            rnorm = sqrt(tmp[0]);
            if (!(rnorm == rnorm)) TOB_FAIL("PCG produced NaN (bad density field)");
            if (rnorm <= tol || rnorm > o.dtol * bnorm) break;
        }
        h->prof_stride = stride_saved1;
        while (!graphed1 && it < o.maxits) {
            TOB_CHECK(elast_apply(h, h->P, h->Q, EPI_DOT, false, 0, nullptr, &np));
            TOB_CHECK(pcg_update_xr_f(c, g, h->U, h->R, h->P, h->Q, h->DINV32, c->d_partials, np, blk_cur, blk_nxt, &nbx));
            it++;
            if (it % check == 0 || it == o.maxits) {
                TOB_CHECK(finalize_parts(c, blk_nxt, nbx, S_TMP));
                TOB_CHECK(read_scalars(c, S_TMP, 2, tmp));
                rnorm = sqrt(tmp[0]);
                if (!(rnorm == rnorm)) TOB_FAIL("PCG produced NaN (bad density field)");
                if (rnorm <= tol || rnorm > o.dtol * bnorm) break;
            }
            TOB_CHECK(pcg_update_p_f(c, g, h->P, h->R, h->DINV32, blk_nxt, blk_cur, nullptr, nbx));
            double* t = blk_cur; blk_cur = blk_nxt; blk_nxt = t;
        }
    } else if (rnorm > tol && bnorm > 0) {
        int cur = S_PAIR_A, nxt = S_PAIR_B;
        if (jac) {
            // alpha = 0/1: leaves x,r untouched and produces rr, rz = r.dinv.r
            TOB_CHECK(pcg_update_xr(c, g, h->U, h->R, h->R, h->R, h->DINV, S_ZERO, S_ONE, cur));
            TOB_CHECK(pcg_update_p(c, g, h->P, h->R, h->DINV, S_ZERO, S_ONE, 1));
        } else {
            TOB_CHECK(gmg_vcycle(h, h->R, h->Z));
            TOB_CHECK(vec_dot(c, g, h->R, h->Z, cur + 1));
            TOB_CHECK(pcg_update_p(c, g, h->P, h->Z, nullptr, S_ZERO, S_ONE, 1));
        }
        const int check = o.check_every > 0 ? o.check_every : 10;
        // Multi-GPU Jacobi-PCG: an iteration is 5 short launches (halo, SpMV, finalize+all-reduce, x/r update,
        // finalize+all-reduce, p update) whose CPU launch cost and dependency gaps exceed their run time on thin
        // slabs.  Blocks of `check` iterations (even, so the scalar-slot roles return to their start) are captured
        // once into a CUDA graph and replayed; two eager iterations per round keep the sampled SpMV event timing.
        const bool graphed = jac && c->desc.nranks > 1 && !getenv("TOB_PCG_NOGRAPH") && check >= 4;
        const int stride_saved = h->prof_stride;
        if (graphed) {
            h->prof_stride = 1;  // the two eager iterations of every round are the timed sample
            const int blk = check & ~1;
            auto one_iteration = [&](int cu, int nx) -> int {
                TOB_CHECK(halo_nodes(c, g, h->P));
                TOB_CHECK(elast_apply(h, h->P, h->Q, EPI_DOT, false, 0, nullptr, &np));
                TOB_CHECK(finalize_sum(c, np, 1, S_PQ, true));
                TOB_CHECK(pcg_update_xr(c, g, h->U, h->R, h->P, h->Q, h->DINV, cu + 1, S_PQ, nx));
                TOB_CHECK(pcg_update_p(c, g, h->P, h->R, h->DINV, nx + 1, cu + 1, 0));
                return 0;
            };
            while (it < o.maxits) {
                // two eager iterations (event-timed when profiling), slots A -> B -> A
                TOB_CHECK(one_iteration(S_PAIR_A, S_PAIR_B));
                TOB_CHECK(one_iteration(S_PAIR_B, S_PAIR_A));
                it += 2;
                if (!h->pcg_graph || h->pcg_graph_its != blk) {
                    if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }
                    const bool prof = h->profile;
                    h->profile = false;  // no event records inside the capture
                    const long long l0 = c->launches;
                    TOB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                    int rc = 0;
                    for (int k = 0; k < blk && !rc; k += 2) {
                        rc = one_iteration(S_PAIR_A, S_PAIR_B);
                        if (!rc) rc = one_iteration(S_PAIR_B, S_PAIR_A);
                    }
                    cudaGraph_t graph = nullptr;
                    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
                    h->profile = prof;
                    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                    if (e != cudaSuccess) TOB_FAIL(std::string("PCG graph capture failed: ") + cudaGetErrorString(e));
                    TOB_CUDA(cudaGraphInstantiate(&h->pcg_graph, graph, 0));
                    cudaGraphDestroy(graph);
                    h->pcg_graph_its = blk;
                    h->pcg_graph_launches = c->launches - l0;
                    c->launches = l0;
                }
                TOB_CUDA(cudaGraphLaunch(h->pcg_graph, c->stream));
                c->launches += h->pcg_graph_launches;
                h->stats.spmv_calls += blk;
                it += blk;
                TOB_CHECK(read_scalars(c, S_PAIR_A, 2, tmp));
                rnorm = sqrt(tmp[0]);
                if (!(rnorm == rnorm)) TOB_FAIL("PCG produced NaN (bad density field)");
                if (rnorm <= tol || rnorm > o.dtol * bnorm) break;
            }
            h->prof_stride = stride_saved;
        }
        while (!graphed && it < o.maxits) {
            TOB_CHECK(halo_nodes(c, g, h->P));
            TOB_CHECK(elast_apply(h, h->P, h->Q, EPI_DOT, false, 0, nullptr, &np));
            TOB_CHECK(finalize_sum(c, np, 1, S_PQ, true));
            // x += a p, r -= a q, [rr, rz_new] -> nxt pair    (alpha = rz_cur / pq)
            TOB_CHECK(pcg_update_xr(c, g, h->U, h->R, h->P, h->Q, jac ? h->DINV : nullptr, cur + 1, S_PQ, nxt));
            it++;
            if (!jac) {
                TOB_CHECK(gmg_vcycle(h, h->R, h->Z));
                TOB_CHECK(vec_dot(c, g, h->R, h->Z, nxt + 1));
            }
            if (it % check == 0 || it == o.maxits || !jac) {
                TOB_CHECK(read_scalars(c, nxt, 2, tmp));
                rnorm = sqrt(tmp[0]);
                if (!(rnorm == rnorm)) TOB_FAIL("PCG produced NaN (indefinite preconditioner or bad density field)");
                if (rnorm <= tol) break;
                if (rnorm > o.dtol * bnorm) break;  // KSP_DIVERGED_DTOL: the reference never checks the reason (LinearElasticity.cc:204-206)
            }
            TOB_CHECK(pcg_update_p(c, g, h->P, jac ? h->R : h->Z, jac ? h->DINV : nullptr, nxt + 1, cur + 1, 0));
            int t = cur; cur = nxt; nxt = t;
        }
    }
    TOB_CUDA(cudaEventRecord(c->ev1, c->stream));
    TOB_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0;
    TOB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    h->stats.iterations = it;
    h->stats.rel_residual = bnorm > 0 ? rnorm / bnorm : 0.0;
    h->stats.seconds = ms * 1e-3;
    h->stats.spmv_seconds = 0.0;
    h->stats.spmv_timed = 0;
    if (h->profile) TOB_CHECK(elast_profile_collect(h));
    if (o.verbose && c->desc.rank == 0)
        printf("State solver:  iter: %i, rerr.: %e, time: %f\n", it, h->stats.rel_residual, h->stats.seconds);
    return 0;
}

static int build_cantilever(tob_elast_impl* h) {
    // N and RHS of LinearElasticity.cc:143-171 on the owned nodes (natural order)
    const LevelGrid& g = *h->g;
    const long long nn = g.owned_nodes();
    std::vector<unsigned char> fixed(3 * nn, 0);
    std::vector<double> rhs(3 * nn, 0.0);
    const double load = -0.001;
    for (int k = g.kn0; k < g.kn1; k++)
        for (int j = 0; j < g.nyn; j++) {
            const long long base = (long long)g.nxn * (j + (long long)g.nyn * (k - g.kn0));
            for (int cc = 0; cc < 3; cc++) fixed[3 * (base + 0) + cc] = 1;  // wall x = xmin clamped (:153-157)
            if (k == 0) {  // line load on the edge x = xmax, z = zmin (:159-170)
                rhs[3 * (base + g.nxn - 1) + 2] = (j == 0 || j == g.nyn - 1) ? load / 2.0 : load;
            }
        }
    return elast_set_bc_load(h, fixed.data(), rhs.data());
}

int elast_set_bc_load(tob_elast_impl* h, const unsigned char* fixed, const double* rhs) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    const long long nd = 3 * g.owned_nodes();
    const void* dfix; const void* drhs;
    TOB_CHECK(stage_in(c, 0, fixed, (size_t)nd, &dfix));
    TOB_CHECK(stage_in(c, 1, rhs, sizeof(double) * nd, &drhs));
    TOB_CUDA(cudaMemsetAsync(h->mask, 0, (size_t)g.mstride * g.nzl(), c->stream));
    TOB_CHECK(nat_to_stored_mask(c, g, (const unsigned char*)dfix, h->mask));
    // ghost planes of the mask: exchange as bytes through a double-typed view is awkward; rebuild by
    // sending the plane through NCCL as chars
    if (c->desc.nranks > 1) {
        const int r = c->desc.rank, n = c->desc.nranks;
        const size_t cnt = (size_t)g.mstride;
        TOB_NCCL(ncclGroupStart());
        if (r > 0) {
            TOB_NCCL(ncclSend(h->mask + (long long)g.zl_own0() * g.mstride, cnt, ncclUint8, r - 1, c->comm, c->stream));
            TOB_NCCL(ncclRecv(h->mask, cnt, ncclUint8, r - 1, c->comm, c->stream));
        }
        if (r < n - 1) {
            TOB_NCCL(ncclSend(h->mask + (long long)(g.zl_own1() - 1) * g.mstride, cnt, ncclUint8, r + 1, c->comm, c->stream));
            TOB_NCCL(ncclRecv(h->mask + (long long)(g.nzl() - 1) * g.mstride, cnt, ncclUint8, r + 1, c->comm, c->stream));
        }
        TOB_NCCL(ncclGroupEnd());
    }
    TOB_CHECK(vec_zero(c, h->RHS, g.nvec));
    TOB_CHECK(nat_to_stored_dof(c, g, (const double*)drhs, h->RHS));
    // NFREE = 1 on free dofs, 0 on clamped (the reference's Vec N); B = N.RHS (LinearElasticity.cc:542)
    TOB_CHECK(mask_to_free(c, g, h->mask, h->NFREE));
    TOB_CHECK(vec_pointwise_mult(c, h->B, h->RHS, h->NFREE, g.nvec));
    TOB_CHECK(vec_pointwise_mult(c, h->U, h->U, h->NFREE, g.nvec));
    TOB_CHECK(vec_dot(c, g, h->B, h->B, S_BB));
    double bb = 0;
    TOB_CHECK(read_scalars(c, S_BB, 1, &bb));
    h->bnorm = sqrt(bb);
    if (h->have_density) {
        TOB_CHECK(hex8_diag_inv(c, g, h->op, h->E_st, h->mask, h->DINV, true));
        TOB_CHECK(vec_d2f(c, h->DINV, h->DINV32, g.nvec));
    }
    if (h->mg) {
        TOB_CHECK(gmg_refresh_masks(h));
        if (h->have_density) TOB_CHECK(gmg_update(h));
    }
    return 0;
}

__global__ void k_mask_to_free(const unsigned char* __restrict__ mask, double* __restrict__ nf, int nxn, int nyn, int nzl,
                               int px, int pm, long long nstride, long long pstride, long long mstride) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)nxn * nyn * nzl;
    if (idx >= tot) return;
    int i = idx % nxn; long long t = idx / nxn;
    int j = t % nyn; int k = t / nyn;
    const unsigned char m = mask[k * mstride + (long long)j * pm + i];
    for (int cc = 0; cc < 3; cc++) nf[k * pstride + cc * nstride + (long long)j * px + i] = (m & (1 << cc)) ? 0.0 : 1.0;
}
int mask_to_free(tob_ctx_impl* c, const LevelGrid& g, const unsigned char* mask, double* nf) {
    TOB_CHECK(vec_zero(c, nf, g.nvec));
    long long tot = (long long)g.nxn * g.nyn * g.nzl();
    k_mask_to_free<<<div_up(tot, 256), 256, 0, c->stream>>>(mask, nf, g.nxn, g.nyn, g.nzl(), g.px, g.pm, g.nstride,
                                                           g.pstride, g.mstride);
    c->launches++;
    TOB_KERNEL_CHECK();
    return 0;
}

static int alloc_vec(tob_ctx_impl* c, double** p, long long n) {
    TOB_CUDA(cudaMalloc(p, sizeof(double) * n));
    TOB_CUDA(cudaMemsetAsync(*p, 0, sizeof(double) * n, c->stream));
    return 0;
}

int elast_create(tob_ctx_impl* c, double nu, const tob_solver_opts* opts, tob_elast_impl** out) {
    auto* h = new tob_elast_impl();
    h->c = c;
    h->g = &c->g0;
    h->nu = nu;
    if (opts) h->opts = *opts; else tob_solver_opts_default(&h->opts);
    const LevelGrid& g = *h->g;
    TOB_CUDA(cudaSetDevice(c->device));
    TOB_CHECK(init_const_slots(c));
    TOB_CHECK(hex8_build_operator(g.dx, g.dy, g.dz, nu, h->op));
    double** vecs[] = {&h->U, &h->B, &h->RHS, &h->R, &h->P, &h->Q, &h->Z, &h->DINV, &h->NFREE};
    for (auto v : vecs) TOB_CHECK(alloc_vec(c, v, g.nvec));
    TOB_CHECK(alloc_vec(c, &h->xphys_st, g.evec));
    TOB_CHECK(alloc_vec(c, &h->E_st, g.evec));
    TOB_CUDA(cudaMalloc(&h->mask, (size_t)g.mstride * g.nzl()));
    TOB_CUDA(cudaMalloc(&h->DINV32, sizeof(float) * g.nvec));
    TOB_CUDA(cudaMemsetAsync(h->DINV32, 0, sizeof(float) * g.nvec, c->stream));
    TOB_CHECK(alloc_vec(c, &h->nat_dof, 3 * g.owned_nodes()));
    TOB_CHECK(alloc_vec(c, &h->nat_dof2, 3 * g.owned_nodes()));
    TOB_CHECK(alloc_vec(c, &h->nat_el, g.owned_elems()));
    TOB_CHECK(alloc_vec(c, &h->nat_el2, g.owned_elems()));
    TOB_CHECK(build_cantilever(h));
    if (h->opts.pc == 1) TOB_CHECK(gmg_setup(h));
    *out = h;
    return 0;
}

void elast_destroy(tob_elast_impl* h) {
    if (!h) return;
    cudaSetDevice(h->c->device);
    cudaStreamSynchronize(h->c->stream);
    gmg_destroy(h);
    double* vecs[] = {h->U, h->B, h->RHS, h->R, h->P, h->Q, h->Z, h->DINV, h->NFREE, h->xphys_st, h->E_st,
                      h->nat_dof, h->nat_dof2, h->nat_el, h->nat_el2};
    for (auto v : vecs) cudaFree(v);
    for (auto e : h->prof_ev) cudaEventDestroy(e);
    if (h->pcg_graph) cudaGraphExecDestroy(h->pcg_graph);
    sr_destroy(h);
    cudaFree(h->DINV32);
    cudaFree(h->u32); cudaFree(h->y32); cudaFree(h->keT32); cudaFree(h->x4); cudaFree(h->y4); cudaFree(h->E32); cudaFree(h->tcB); cudaFree(h->tcB3);
    cudaFree(h->mask);
    hex8_free_operator(h->op);
    delete h;
}

}  // namespace tob

using namespace tob;

extern "C" {

void tob_solver_opts_default(tob_solver_opts* o) {
    o->pc = 0;
    o->nlvls = 4;       // TopOpt.cc:116
    o->rtol = 1.0e-5;   // LinearElasticity.cc:621
    o->atol = 1.0e-50;  // :622
    o->dtol = 1.0e5;    // :623
    o->maxits = 20000;  // Jacobi-PCG needs far more than the reference's 200 FGMRES+MG iterations
    o->check_every = 25;
    o->cheb_degree = 4;  // reference smoother: 4 GMRES+SOR iterations (:734-746)
    o->coarse_its = 30;  // reference coarse solve: GMRES(30) (:720-731)
    o->spmv_kernel = 0;
    o->verbose = 0;
}

int tob_elast_create(tob_ctx* ctx, double nu, const tob_solver_opts* opts, tob_elast** out) {
    if (!ctx || !out) TOB_FAIL("null argument");
    tob_elast_impl* h = nullptr;
    TOB_CHECK(elast_create(reinterpret_cast<tob_ctx_impl*>(ctx), nu, opts, &h));
    *out = reinterpret_cast<tob_elast*>(h);
    return 0;
}
int tob_elast_destroy(tob_elast* h) {
    elast_destroy(reinterpret_cast<tob_elast_impl*>(h));
    return 0;
}
int tob_elast_set_opts(tob_elast* hh, const tob_solver_opts* o) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    const bool need_mg = o->pc == 1 && !h->mg;
    if (h->mg) gmg_drop_graph(h);
    if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }
    if (h->sr_graph) { cudaGraphExecDestroy(h->sr_graph); h->sr_graph = nullptr; }
    h->opts = *o;
    if (need_mg) {
        TOB_CHECK(gmg_setup(h));
        if (h->have_density) TOB_CHECK(gmg_update(h));
    }
    return 0;
}
int tob_elast_set_bc_load(tob_elast* hh, const unsigned char* fixed, const double* rhs) {
    return elast_set_bc_load(reinterpret_cast<tob_elast_impl*>(hh), fixed, rhs);
}
int tob_elast_get_ke(tob_elast* hh, double* ke) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    memcpy(ke, h->op.KE, sizeof(double) * 576);
    return 0;
}
int tob_elast_set_density(tob_elast* hh, const double* xPhys, double Emin, double Emax, double penal) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    const void* dx;
    TOB_CHECK(stage_in(h->c, 0, xPhys, sizeof(double) * h->g->owned_elems(), &dx));
    return elast_set_density(h, (const double*)dx, Emin, Emax, penal);
}
int tob_elast_spmv(tob_elast* hh, const double* u, double* y) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    if (!h->have_density) TOB_FAIL("tob_elast_set_density must be called before tob_elast_spmv");
    const long long nd = 3 * g.owned_nodes();
    const void* du; void* dy;
    TOB_CHECK(stage_in(c, 0, u, sizeof(double) * nd, &du));
    TOB_CHECK(stage_out_begin(c, 1, y, sizeof(double) * nd, &dy));
    TOB_CHECK(vec_zero(c, h->P, g.nvec));
    TOB_CHECK(nat_to_stored_dof(c, g, (const double*)du, h->P));
    TOB_CHECK(halo_nodes(c, g, h->P));
    TOB_CHECK(elast_apply(h, h->P, h->Q, EPI_Y, true, 0, nullptr, nullptr));
    TOB_CHECK(stored_to_nat_dof(c, g, h->Q, (double*)dy));
    return stage_out_finish(c, y, dy, sizeof(double) * nd);
}
static int tc32_buffers(tob_elast_impl* h) {
    if (h->u32) return 0;
    TOB_CUDA(cudaMalloc(&h->u32, sizeof(float) * h->g->nvec));
    TOB_CUDA(cudaMalloc(&h->y32, sizeof(float) * h->g->nvec));
    TOB_CUDA(cudaMemsetAsync(h->y32, 0, sizeof(float) * h->g->nvec, h->c->stream));
    return 0;
}
}  // extern "C"
// buffers / operand blocks / fp32 modulus of the generation-2 tensor-core operator
int tob::tc2_prepare(tob_elast_impl* h) {
    const LevelGrid& g = *h->g;
    const long long nn = v4_plane_nodes(g) * g.nzl();
    if (!h->x4) {
        TOB_CUDA(cudaMalloc(&h->x4, sizeof(float4) * nn));
        TOB_CUDA(cudaMalloc(&h->y4, sizeof(float4) * nn));
        TOB_CUDA(cudaMemsetAsync(h->y4, 0, sizeof(float4) * nn, h->c->stream));
        TOB_CUDA(cudaMalloc(&h->E32, sizeof(float) * g.evec));
        TOB_CHECK(tc2_build_B(h->op, &h->tcB));
        TOB_CHECK(tc3_build_B(h->op, &h->tcB3));
    }
    if (!h->E32_valid) {
        TOB_CHECK(vec_d2f(h->c, h->E_st, h->E32, g.evec));
        h->E32_valid = true;
    }
    return 0;
}
extern "C" {
/* y = N.(K (N.u)) + (1-N).u in single precision on the tensor cores (TF32 inputs, fp32 accumulation): the smoother-grade
 * variant of tob_elast_spmv.  u / y are fp64 at the boundary (converted inside). */
int tob_elast_spmv_tf32(tob_elast* hh, const double* u, double* y) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    if (!h->have_density) TOB_FAIL("tob_elast_set_density must be called before tob_elast_spmv_tf32");
    const long long nd = 3 * g.owned_nodes();
    const void* du; void* dy;
    TOB_CHECK(stage_in(c, 0, u, sizeof(double) * nd, &du));
    TOB_CHECK(stage_out_begin(c, 1, y, sizeof(double) * nd, &dy));
    TOB_CHECK(tc32_buffers(h));
    TOB_CHECK(vec_zero(c, h->P, g.nvec));
    TOB_CHECK(nat_to_stored_dof(c, g, (const double*)du, h->P));
    TOB_CHECK(halo_nodes(c, g, h->P));
    const int gen = getenv("TOB_TC32_GEN") ? atoi(getenv("TOB_TC32_GEN")) : 3;
    if (gen == 1) {
        TOB_CHECK(vec_d2f(c, h->P, h->u32, g.nvec));
        TOB_CHECK(hex8_spmv_tc32(c, g, h->op, h->u32, h->y32, h->E_st, h->mask, &h->keT32));
        TOB_CHECK(vec_f2d(c, h->y32, h->Q, g.nvec));
    } else {
        // generation 2 works on pre-masked node-interleaved fp32 vectors (the smoother's own layout)
        TOB_CHECK(tc2_prepare(h));
        TOB_CHECK(vec_pointwise_mult(c, h->Z, h->P, h->NFREE, g.nvec));
        TOB_CHECK(vec_st_to_v4(c, g, h->Z, h->x4));
        if (gen == 2) TOB_CHECK(hex8_tc32b(c, g, h->tcB, h->x4, h->y4, h->E32, h->mask, 0, nullptr, nullptr, nullptr, 0.f, 0.f));
        else TOB_CHECK(hex8_tc32c(c, g, h->tcB3, h->x4, h->y4, h->E32, h->mask, 0, nullptr, nullptr, nullptr, 0.f, 0.f));
        TOB_CHECK(vec_v4_to_st(c, g, h->y4, h->Q));
        TOB_CHECK(vec_restore_fixed(c, h->Q + g.own_off, h->P + g.own_off, h->NFREE + g.own_off, g.own_len));
    }
    TOB_CHECK(stored_to_nat_dof(c, g, h->Q, (double*)dy));
    return stage_out_finish(c, y, dy, sizeof(double) * nd);
}
/* times `reps` launches of the tensor-core operator (L2 flushed between launches, CUDA events around each) */
int tob_elast_spmv_tf32_bench(tob_elast* hh, int reps, double* ms_per_apply) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const int gen = getenv("TOB_TC32_GEN") ? atoi(getenv("TOB_TC32_GEN")) : 3;
    if (!h->have_density || (gen == 1 ? !h->u32 : !h->x4)) TOB_FAIL("call tob_elast_spmv_tf32 once first");
    const size_t flush_bytes = 512ull << 20;
    TOB_CHECK(c->stage[7].reserve(flush_bytes));
    auto apply = [&]() -> int {
        if (gen == 1) return hex8_spmv_tc32(c, *h->g, h->op, h->u32, h->y32, h->E_st, h->mask, &h->keT32);
        if (gen == 2) return hex8_tc32b(c, *h->g, h->tcB, h->x4, h->y4, h->E32, h->mask, 0, nullptr, nullptr, nullptr, 0.f, 0.f);
        return hex8_tc32c(c, *h->g, h->tcB3, h->x4, h->y4, h->E32, h->mask, 0, nullptr, nullptr, nullptr, 0.f, 0.f);
    };
    for (int i = 0; i < 3; i++) TOB_CHECK(apply());
    cudaEvent_t e0, e1;
    TOB_CUDA(cudaEventCreate(&e0));
    TOB_CUDA(cudaEventCreate(&e1));
    double tot = 0.0;
    for (int i = 0; i < reps; i++) {
        TOB_CUDA(cudaMemsetAsync(c->stage[7].p, i & 0xff, flush_bytes, c->stream));
        TOB_CUDA(cudaEventRecord(e0, c->stream));
        TOB_CHECK(apply());
        TOB_CUDA(cudaEventRecord(e1, c->stream));
        TOB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        TOB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        tot += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_apply = reps ? tot / reps : 0.0;
    return 0;
}
int tob_elast_diag(tob_elast* hh, double* d) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    if (!h->have_density) TOB_FAIL("tob_elast_set_density must be called first");
    const long long nd = 3 * g.owned_nodes();
    void* dd;
    TOB_CHECK(stage_out_begin(c, 1, d, sizeof(double) * nd, &dd));
    TOB_CHECK(hex8_diag_inv(c, g, h->op, h->E_st, h->mask, h->Q, false));
    TOB_CHECK(stored_to_nat_dof(c, g, h->Q, (double*)dd));
    return stage_out_finish(c, d, dd, sizeof(double) * nd);
}
int tob_elast_spmv_bench(tob_elast* hh, int kernel, int reps, double* ms_per_apply) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    tob_ctx_impl* c = h->c;
    if (!h->have_density) TOB_FAIL("tob_elast_set_density must be called first");
    // P holds whatever the last call left (tests fill it through tob_elast_spmv).  Between timed
    // applications a 512 MB scratch buffer is overwritten so that every application starts with a cold
    // L2 (126 MB); only the SpMV launches themselves are inside the event pairs.
    const size_t flush_bytes = 512ull << 20;
    TOB_CHECK(c->stage[7].reserve(flush_bytes));
    for (int i = 0; i < 3; i++) TOB_CHECK(elast_apply(h, h->P, h->Q, EPI_Y, false, kernel, nullptr, nullptr));
    const bool prof = h->profile;
    h->profile = true;
    h->prof_used = 0;
    h->stats.spmv_seconds = 0.0;
    h->stats.spmv_timed = 0;
    for (int i = 0; i < reps; i++) {
        TOB_CUDA(cudaMemsetAsync(c->stage[7].p, i & 0xff, flush_bytes, c->stream));
        TOB_CHECK(elast_apply(h, h->P, h->Q, EPI_Y, false, kernel, nullptr, nullptr));
    }
    TOB_CUDA(cudaStreamSynchronize(c->stream));
    TOB_CHECK(elast_profile_collect(h));
    h->profile = prof;
    *ms_per_apply = h->stats.spmv_timed ? h->stats.spmv_seconds * 1e3 / h->stats.spmv_timed : 0.0;
    return 0;
}
int tob_elast_solve(tob_elast* hh, const double* xPhys, double Emin, double Emax, double penal) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    TOB_CHECK(tob_elast_set_density(hh, xPhys, Emin, Emax, penal));
    return elast_pcg(h);
}

static int obj_sens_common(tob_elast_impl* h, double volfrac, double* fx, double* gx, double* dfdx, double* dgdx) {
    tob_ctx_impl* c = h->c;
    const LevelGrid& g = *h->g;
    const long long ne = g.owned_elems();
    const long long nel_glob = (long long)g.nex * g.ney * g.nez;
    void* ddf = h->nat_el;
    if (dfdx) TOB_CHECK(stage_out_begin(c, 1, dfdx, sizeof(double) * ne, &ddf));
    TOB_CHECK(halo_nodes(c, g, h->U));
    TOB_CHECK(hex8_compliance(c, g, h->op, h->U, h->xphys_st, h->E_st, h->penal, h->Emin, h->Emax, (double*)ddf, S_OBJ));
    double s[2];
    TOB_CHECK(read_scalars(c, S_OBJ, 2, s));
    if (fx) *fx = s[0];                              // LinearElasticity.cc:421,427-429
    if (gx) *gx = s[1] / (double)nel_glob - volfrac;  // :432-436
    if (dfdx) TOB_CHECK(stage_out_finish(c, dfdx, ddf, sizeof(double) * ne));
    if (dgdx) {  // VecSet(dgdx, 1/neltot) (:437)
        void* ddg;
        TOB_CHECK(stage_out_begin(c, 2, dgdx, sizeof(double) * ne, &ddg));
        TOB_CHECK(vec_fill(c, (double*)ddg, 1.0 / (double)nel_glob, ne));
        TOB_CHECK(stage_out_finish(c, dgdx, ddg, sizeof(double) * ne));
    }
    return 0;
}

int tob_elast_solve_obj_sens(tob_elast* hh, const double* xPhys, double Emin, double Emax, double penal, double volfrac,
                             double* fx, double* gx, double* dfdx, double* dgdx) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    TOB_CHECK(tob_elast_solve(hh, xPhys, Emin, Emax, penal));
    return obj_sens_common(h, volfrac, fx, gx, dfdx, dgdx);
}
int tob_elast_obj(tob_elast* hh, const double* xPhys, double Emin, double Emax, double penal, double volfrac,
                  double* fx, double* gx) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    TOB_CHECK(tob_elast_solve(hh, xPhys, Emin, Emax, penal));
    return obj_sens_common(h, volfrac, fx, gx, nullptr, nullptr);
}
int tob_elast_sens(tob_elast* hh, const double* xPhys, double Emin, double Emax, double penal, double volfrac,
                   double* dfdx, double* dgdx) {
    // ComputeSensitivities (LinearElasticity.cc:299-361) reuses the state of the last solve
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    TOB_CHECK(tob_elast_set_density(hh, xPhys, Emin, Emax, penal));
    return obj_sens_common(h, volfrac, nullptr, nullptr, dfdx, dgdx);
}
int tob_elast_get_state(tob_elast* hh, double* U) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    const long long nd = 3 * h->g->owned_nodes();
    void* du;
    TOB_CHECK(stage_out_begin(h->c, 1, U, sizeof(double) * nd, &du));
    TOB_CHECK(stored_to_nat_dof(h->c, *h->g, h->U, (double*)du));
    return stage_out_finish(h->c, U, du, sizeof(double) * nd);
}
int tob_elast_set_state(tob_elast* hh, const double* U) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    const long long nd = 3 * h->g->owned_nodes();
    const void* du;
    TOB_CHECK(stage_in(h->c, 0, U, sizeof(double) * nd, &du));
    TOB_CHECK(vec_zero(h->c, h->U, h->g->nvec));
    TOB_CHECK(nat_to_stored_dof(h->c, *h->g, (const double*)du, h->U));
    return vec_pointwise_mult(h->c, h->U, h->U, h->NFREE, h->g->nvec);
}
int tob_elast_get_rhs(tob_elast* hh, double* rhs) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    const long long nd = 3 * h->g->owned_nodes();
    void* du;
    TOB_CHECK(stage_out_begin(h->c, 1, rhs, sizeof(double) * nd, &du));
    TOB_CHECK(stored_to_nat_dof(h->c, *h->g, h->RHS, (double*)du));
    return stage_out_finish(h->c, rhs, du, sizeof(double) * nd);
}
int tob_elast_set_profile(tob_elast* hh, int on) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    h->profile = on != 0;
    h->prof_stride = on > 1 ? on : 1;  // on = k > 1: time every k-th fine-level SpMV launch
    h->prof_used = 0;
    return 0;
}
int tob_elast_last_stats(tob_elast* hh, tob_solve_stats* s) {
    auto* h = reinterpret_cast<tob_elast_impl*>(hh);
    *s = h->stats;
    s->kernel_launches = (int)h->c->launches;
    return 0;
}

}  // extern "C"
