#!/usr/bin/env python
"""bench.py -- design iterations/sec (and matrix-free SpMV GB/s) of the B200-native hot path.

A "step" is one design iteration of main.cc:54-105 (state solve + objective/sensitivities, sensitivity filtering, MMA
update, density filter, MND).  Default workload = BASELINE.json configs[1]: synthetic cantilever 192x96x96, matrix-free
PCG + Jacobi (rtol 1e-5, warm start), density filter rmin = 3 cells.

  python bench.py --gpus 1 --steps K --warmup W              (our arm; N>1 under torchrun: z-slabs, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W      (reference arm: the reference's algorithm on the host cores)
  python bench.py --workload cfg3|cfg4|cfg5                  (the other BASELINE configs as the main workload)

Keys of the JSON line (one line, rank 0):
`value`   device-resident design iterations/sec (inputs already in HBM), CUDA events recorded on the library's stream,
          max over ranks.
`e2e`     the same loop driven through the C-ABI with HOST (pinned) buffers -- every call stages its vectors
          host<->device inside the timed region, exactly what the reference's main.cc/MMA.cc do through VecGetArray.
`roofline` the fine-level matrix-free SpMV kernel, timed live with CUDA event pairs around every 8th launch of the eager
          part of the timed region; algorithmic bytes = 8*nel + 16*ndof per application.
`spmv_256` (N = 1) the same kernel alone on 256^3 elements, L2 flushed between launches: the north-star SpMV figure.
`gmg_solver` the cfg2 design iterations with the multigrid-preconditioned solver.
`cfg3_gmg` BASELINE configs[2] (384x192x192, multigrid) on the same N GPUs -- the strong-scaling curve of the 8-GPU job.
`cpu_baseline` (N = 1) ONE WHOLE design iteration of the reference's algorithm on the host cores
          (oracle/cpu_design_loop.py: CSR assembly + Galerkin set-up + FGMRES/PCMG/GMRES-SOR solve + filters + MMA).
Reference arm: whole design iterations of the same CPU code, as many of the requested steps as fit in the time budget
(TOB_REF_BUDGET_S, default 150 s); `steps` is the number actually run and timed -- nothing is extrapolated.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

BASE = dict(volfrac=0.12, penal=3.0, Emin=1e-9, Emax=1.0, Xmin=0.0, Xmax=1.0, movlim=0.2, rtol=1e-5, nu=0.3,
            beta=0.1, beta_final=48.0, eta=0.0)
WORKLOADS = {
    # BASELINE.json configs[1]
    "cfg2": dict(nex=192, ney=96, nez=96, xc=(0, 2, 0, 1, 0, 1), bc="cantilever", filter=1, rmin_cells=3.0, proj=False,
                 solver="jacobi",
                 desc="cantilever 192x96x96, matrix-free PCG+Jacobi rtol 1e-5 warm-started, density filter rmin=3 cells, MMA m=1"),
    # configs[2]
    "cfg3": dict(nex=384, ney=192, nez=192, xc=(0, 2, 0, 1, 0, 1), bc="cantilever", filter=1, rmin_cells=3.0, proj=False,
                 solver="gmg",
                 desc="cantilever 384x192x192, matrix-free PCG + geometric multigrid V-cycle, density filter rmin=3 cells, MMA m=1"),
    # configs[3]: half of a 2:1:1 MBB beam (symmetry plane x = 0), see mbb_bc_load
    "cfg4": dict(nex=256, ney=256, nez=256, xc=(0, 1, 0, 1, 0, 1), bc="mbb", filter=2, rmin_cells=6.0, proj=True,
                 solver="gmg",
                 desc="MBB half-beam 256^3 (symmetry plane x=0, roller edge, line load), PDE (Helmholtz) filter rmin=6 cells + "
                      "Heaviside projection with beta continuation 0.1->48, matrix-free PCG + multigrid, MMA m=1"),
    # configs[4]
    "cfg5": dict(nex=1024, ney=512, nez=512, xc=(0, 2, 0, 1, 0, 1), bc="bridge", filter=1, rmin_cells=3.0, proj=False,
                 solver="gmg",
                 desc="bridge 1024x512x512 (clamped bottom edges x=xmin/xmax, deck load on z=zmax), fp64, matrix-free PCG + "
                      "multigrid, density filter rmin=3 cells, MMA m=1"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


# ---- BC / load definitions (LinearElasticity.cc:145-171 is the template; MBB and bridge are not in the reference) --------
def bc_load(kind, nxn, nyn, nzn, kn0, kn1):
    """fixed (uint8) and rhs (float64) for this rank's owned node planes [kn0, kn1), natural order, 3 dofs per node.
    Loads are per node with half weight at the two y-end nodes of a loaded line, as the reference does (:159-170)."""
    nz = kn1 - kn0
    fixed = np.zeros((nz, nyn, nxn, 3), dtype=np.uint8)
    rhs = np.zeros((nz, nyn, nxn, 3))
    yw = np.ones(nyn)
    yw[0] = yw[-1] = 0.5
    ks = np.arange(kn0, kn1)
    if kind == "cantilever":  # the reference: wall x = xmin clamped, line load on the edge x = xmax, z = zmin
        fixed[:, :, 0, :] = 1
        if kn0 == 0:
            rhs[0, :, nxn - 1, 2] = -0.001 * yw
    elif kind == "mbb":
        # half MBB beam: symmetry plane x = xmin (u_x = 0); roller on the bottom edge x = xmax, z = zmin (u_z = 0, and
        # u_y = 0 there to remove the rigid translation along y); line load -0.001 per node on the top edge x = xmin, z = zmax
        fixed[:, :, 0, 0] = 1
        if kn0 == 0:
            fixed[0, :, nxn - 1, 2] = 1
            fixed[0, :, nxn - 1, 1] = 1
        sel = ks == nzn - 1
        if sel.any():
            rhs[np.where(sel)[0][0], :, 0, 2] = -0.001 * yw
    elif kind == "bridge":
        # clamped bottom edges (x = xmin and x = xmax at z = zmin, all three dofs); deck load on the top face z = zmax:
        # -0.001/nxn per node (half on the boundary lines), so the total force matches the cantilever's line load
        if kn0 == 0:
            fixed[0, :, 0, :] = 1
            fixed[0, :, nxn - 1, :] = 1
        sel = ks == nzn - 1
        if sel.any():
            xw = np.ones(nxn)
            xw[0] = xw[-1] = 0.5
            rhs[np.where(sel)[0][0], :, :, 2] = -0.001 / nxn * np.outer(yw, xw)
    else:
        raise ValueError(kind)
    return fixed.reshape(-1), rhs.reshape(-1)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def gmg_levels(c, cap=8):
    """deepest hierarchy (<= cap levels) whose coarsening divides the grid (on N > 1 ranks the library replicates the
    levels whose slab boundaries stop being aligned, so the depth does not depend on N)"""
    nlv = cap
    # every dimension must divide, and the coarsest grid keeps >= 2 elements per direction (it is inverted explicitly)
    while nlv > 2 and any(d % (1 << (nlv - 1)) or (d >> (nlv - 1)) < 2 for d in (c["nex"], c["ney"], c["nez"])):
        nlv -= 1
    return nlv


class DesignLoop:
    """main.cc:48-105 over the C-ABI.  resident=True keeps every vector in HBM (DevVec); resident=False uses
    pinned host numpy buffers, so each C-ABI call stages host<->device inside the timed region."""

    def __init__(self, ctx, cfg, resident, profile=True, pc=0, nlvls=4):
        from topopt_in_petsc_b200 import Filter, LinearElasticity, MMA, solver_opts
        self.ctx, self.resident, self.c = ctx, resident, cfg
        c = cfg
        self.phys = LinearElasticity(ctx, nu=c["nu"], opts=solver_opts(pc=pc, nlvls=nlvls, rtol=c["rtol"], cheb_degree=3,
                                                                       coarse_its=20 if nlvls >= 5 else 60,
                                                                       maxits=100000 if pc == 0 else 500, check_every=50))
        if c["bc"] != "cantilever":
            fixed, rhs = bc_load(c["bc"], ctx.nxn, ctx.nyn, ctx.nzn, ctx.kn0, ctx.kn1)
            self.phys.SetBCLoad(fixed, rhs)
        # every 8th fine-level SpMV launch is bracketed by CUDA events (timing all of them costs ~6 % of the step)
        self.phys.SetProfile(0 if (not profile or os.environ.get("TOB_BENCH_NOPROFILE")) else 8)
        dx = (c["xc"][1] - c["xc"][0]) / c["nex"]
        self.flt = Filter(ctx, c["filter"], c["rmin_cells"] * dx)
        self.mma = MMA(ctx, 1)
        self.beta = c["beta"]
        n = ctx.nel_owned
        self.n = n
        names = ["x", "xTilde", "xPhys", "dfdx", "dgdx", "xmin", "xmax", "xold"]
        if resident:
            self.v = {k: ctx.dev(n) for k in names}
            for k in ("x", "xTilde", "xPhys", "xold"):
                self.v[k].upload(np.full(n, c["volfrac"]))
        else:
            import torch
            self._pinned = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in names}
            self.v = {k: t.numpy() for k, t in self._pinned.items()}
            for k in ("x", "xTilde", "xPhys", "xold"):
                self.v[k][:] = c["volfrac"]
        self.fscale = None
        self.itr = 0
        self.spmv_s, self.spmv_n, self.spmv_calls, self.pcg_its = 0.0, 0, 0, []
        self.h2d = self.d2h = 0
        self.flt.FilterProject(self.v["x"], self.v["xTilde"], self.v["xPhys"], c["proj"], self.beta, c["eta"])  # main.cc:48
        self.hist = []

    def step(self):
        c, v, n8 = self.c, self.v, 8 * self.n
        self.itr += 1
        fx, gx = self.phys.ComputeObjectiveConstraintsSensitivities(v["dfdx"], v["dgdx"], v["xPhys"], c["Emin"], c["Emax"],
                                                                    c["penal"], c["volfrac"])  # main.cc:62
        st = self.phys.stats()
        self.spmv_s += st.spmv_seconds
        self.spmv_n += st.spmv_timed
        self.spmv_calls += st.iterations + 1  # one fine-level SpMV per Krylov iteration + the initial residual
        self.pcg_its.append(st.iterations)
        if self.itr == 1:
            self.fscale = 10.0 / fx  # main.cc:68-70
        self.ctx.scale(v["dfdx"], self.fscale, self.n)  # VecScale, main.cc:73
        self.flt.Gradients(v["x"], v["xTilde"], v["dfdx"], 1, [v["dgdx"]], c["proj"], self.beta, c["eta"])  # :76
        self.mma.SetOuterMovelimit(c["Xmin"], c["Xmax"], c["movlim"], v["x"], v["xmin"], v["xmax"])  # :81
        self.mma.Update(v["x"], v["dfdx"], np.array([gx]), [v["dgdx"]], v["xmin"], v["xmax"])  # :85
        ch = self.mma.DesignChange(v["x"], v["xold"])  # :89
        if c["proj"]:
            self.beta, _ = self.flt.IncreaseBeta(self.beta, c["beta_final"], gx, self.itr, ch)  # :94
        self.flt.FilterProject(v["x"], v["xTilde"], v["xPhys"], c["proj"], self.beta, c["eta"])  # :98
        mnd = self.flt.GetMND(v["xPhys"])  # :102
        if not self.resident:
            # bytes staged by the calls above (host pointers): inputs H2D / outputs D2H, counted per vector
            self.h2d += n8 * (1 + 1 + 4 + 1 + 5 + 2 + 1 + 1)
            self.d2h += n8 * (2 + 1 + 2 + 2 + 1 + 1 + 2)
        self.hist.append(dict(itr=self.itr, fx=fx, gx=gx, ch=ch, mnd=mnd, pcg_its=st.iterations, beta=self.beta))
        return fx

    def close(self):
        for o in (self.mma, self.flt, self.phys):
            o.close()
        if self.resident:
            for d in self.v.values():
                d.free()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from topopt_in_petsc_b200 import Context, LinearElasticity, nccl_unique_id
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    c = dict(BASE, **WORKLOADS[args.workload])

    def make_ctx(cfg):
        # every communicator needs its own ncclUniqueId
        cid = None
        if world > 1:
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, 0)
            cid = bytes(t.cpu().tolist())
        return Context(cfg["nex"] + 1, cfg["ney"] + 1, cfg["nez"] + 1, xc=cfg["xc"], rank=rank, nranks=world, device=local,
                       nccl_id=cid)

    ctx = make_ctx(c)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(cx, loop, K, W, sampler=None):
        # the library works on its own (non-blocking) stream: the timing events are recorded on THAT stream
        lib_stream = torch.cuda.ExternalStream(cx.stream())
        for _ in range(W):
            loop.step()
        loop.spmv_s, loop.spmv_n, loop.spmv_calls = 0.0, 0, 0
        loop.h2d = loop.d2h = 0
        nwarm = len(loop.pcg_its)
        cx.launch_count(reset=True)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        for _ in range(K):
            loop.step()
        e1.record(lib_stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks, loop.pcg_its[nwarm:]

    K, W = args.steps, args.warmup
    nlv = gmg_levels(c)
    main_pc = 1 if (args.solver == "gmg" or c["solver"] == "gmg") else 0
    # ---- device-resident arm -----------------------------------------------------------------
    loop = DesignLoop(ctx, c, resident=True, pc=main_pc, nlvls=nlv)
    ms, clocks, its = timed(ctx, loop, K, W, ClockSampler(local))
    launches = ctx.launch_count()
    spmv_s, spmv_n, spmv_calls = loop.spmv_s, loop.spmv_n, loop.spmv_calls
    hist = loop.hist
    loop.close()
    # ---- end-to-end arm: host (pinned) buffers through the same C-ABI calls ---------------------------
    loop2 = DesignLoop(ctx, c, resident=False, pc=main_pc, nlvls=nlv)
    ms2, _, _ = timed(ctx, loop2, K, W)
    h2d, d2h = loop2.h2d // K, loop2.d2h // K
    loop2.close()

    # ---- same job with the multigrid-preconditioned solver (not the BASELINE-named cfg2 solver; reported beside it) ----
    gmg = None
    if not args.no_gmg and main_pc == 0:
        try:  # a secondary leg must never take the headline measurement down with it
            loop3 = DesignLoop(ctx, c, resident=True, pc=1, nlvls=nlv)
            ms3, _, its3 = timed(ctx, loop3, K, W)
            gmg = {"value": K / (ms3 * 1e-3), "unit": "design iterations/s (whole job)", "ms_per_step": ms3 / K, "nlvls": nlv,
                   "pcg_iterations_per_step": its3, "smoother": "Chebyshev-Jacobi degree 3",
                   "final_fx": loop3.hist[-1]["fx"]}
            loop3.close()
        except Exception as e:  # noqa: BLE001
            gmg = {"error": str(e)[:300]}
    nel, ndof, nel_loc, ndof_loc = ctx.nel, ctx.ndof, ctx.nel_owned, ctx.ndof_owned
    p2p = bool(int(os.environ.get("TOB_P2P", "1")))
    ctx.close()

    # ---- BASELINE configs[2] on the same GPUs (scaling curve of the 8-GPU job) ---------------------------------------
    cfg3 = None
    if not args.no_cfg3 and args.workload == "cfg2":
        try:
            c3 = dict(BASE, **WORKLOADS["cfg3"])
            ctx3 = make_ctx(c3)
            K3, W3 = max(10, min(K, 12)), 3
            l3 = DesignLoop(ctx3, c3, resident=True, profile=False, pc=1, nlvls=gmg_levels(c3))
            ms3, _, its3 = timed(ctx3, l3, K3, W3)
            cfg3 = {"workload": c3["desc"], "value": K3 / (ms3 * 1e-3), "unit": "design iterations/s (whole job)",
                    "ms_per_step": ms3 / K3, "steps": K3, "warmup": W3, "n_gpus": world, "nlvls": gmg_levels(c3),
                    "pcg_iterations_per_step": its3, "final_fx": l3.hist[-1]["fx"], "scaling": "strong"}
            l3.close()
            ctx3.close()
        except Exception as e:  # noqa: BLE001
            cfg3 = {"error": str(e)[:300]}

    # ---- north-star SpMV figure: the kernel alone at 256^3, L2 flushed between launches (one GPU) -------------------
    spmv256 = None
    if world == 1 and not args.no_spmv256 and args.workload == "cfg2":
        try:
            peak, kind = peaks()
            cx = Context(257, 257, 257, device=local)
            ph = LinearElasticity(cx)
            ph.SetDensity(np.full(cx.nel, 0.12), 1e-9, 1.0, 3.0)
            ph.SpMV(np.random.default_rng(1234).uniform(-1, 1, cx.ndof))  # leaves u resident in the operator's input buffer
            msk = ph.SpMVBench(kernel=0, reps=20)
            alg256 = 8 * cx.nel + 16 * cx.ndof
            tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
            tr = json.load(open(tp)).get("256x256x256", {}).get("dram_bytes_per_launch") if os.path.exists(tp) else None
            spmv256 = {"kernel": "default fine-level SpMV kernel (k_hex8_zm4, fp64)", "elements": [256, 256, 256], "ms": msk,
                       "achieved": alg256 / (msk * 1e-3) / 1e9, "unit": "GB/s", "peak": peak, "peak_kind": kind,
                       "frac": alg256 / (msk * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg256, "traffic": tr,
                       "l2": "512 MB memset between launches (cold L2), CUDA events around each launch, 20 launches"}
            ph.close()
            cx.close()
        except Exception as e:  # noqa: BLE001
            spmv256 = {"error": str(e)[:300]}

    if rank == 0:
        peak, kind = peaks()
        # per-launch algorithmic bytes of THIS rank's slab (SpMV is launched once per rank per application)
        alg = 8 * nel_loc + 16 * ndof_loc
        achieved = alg * spmv_n / spmv_s / 1e9 if spmv_s > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
        if os.path.exists(tp) and world == 1:
            traffic = json.load(open(tp)).get("x".join(str(c[k]) for k in ("nex", "ney", "nez")), {}).get("dram_bytes_per_launch")
        out = {
            "metric": "design_iters_per_sec", "value": K / (ms * 1e-3), "unit": "design iterations/s (whole job)",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": c["desc"], "elements": [c["nex"], c["ney"], c["nez"]], "nel": nel, "ndof": ndof,
                       "partition": f"z-slab x{world}", "solver": "pcg+gmg" if main_pc else "pcg+jacobi",
                       "l2": "solver working set (6 vectors of 8*ndof bytes + modulus) exceeds the 126 MB L2; every SpMV is "
                             "separated by vector-update traffic larger than L2",
                       "pcg_iterations_per_step": its, "timed_design_iterations": [h["itr"] for h in hist[W:]]},
            "clocks": clocks, "gpu_launches": int(launches),
            "comm": {"nranks": world, "transport": ("NVLink peer-memory kernels (CUDA IPC) for halos/all-reduce; NCCL for set-up"
                                                    if (world > 1 and p2p) else ("NCCL" if world > 1 else "none"))},
            "e2e": {"value": K / (ms2 * 1e-3), "unit": "design iterations/s (whole job)", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms2 / K},
            "roofline": {"kernel": "k_hex8_zm4 (fine-level matrix-free hex8 SpMV, fp64), in-solver", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_kind": kind, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg, "launches_timed": spmv_n, "launches_total": spmv_calls,
                         "avg_launch_us": (spmv_s / spmv_n * 1e6) if spmv_n else None,
                         "share_of_step": (spmv_s / spmv_n * spmv_calls / (ms * 1e-3)) if spmv_n else None},
            "spmv_256": spmv256,
            "history": [{k: (round(v, 9) if isinstance(v, float) else v) for k, v in h.items()} for h in hist[W:]],
            "gmg_solver": gmg,
            "cfg3_gmg": cfg3,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_whole_steps(c, steps=1, warmup=0, budget_s=1e9)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def host_cores():
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        return os.cpu_count() or 1


def pick_cpu_threads(c, cores, nlv):
    """Thread count of the CPU arm: the restated solver is memory-bound and its smoothers are block-local (one OpenMP thread = one
    PETSc rank), so "all cores" is not always the fastest choice (a 64-core box ran one FGMRES iteration of cfg2 in 2.6 s with 64
    threads and about half of that with 32).  Time two FGMRES iterations of the actual problem for cores, cores/2, cores/4 and
    keep the fastest; TOB_REF_THREADS overrides.  Any failure falls back to all cores."""
    if os.environ.get("TOB_REF_THREADS"):
        return max(1, int(os.environ["TOB_REF_THREADS"])), {"forced": True}
    cands = sorted({n for n in (cores, cores // 2, cores // 4) if n >= 4}, reverse=True) or [cores]
    if len(cands) == 1:
        return cands[0], {}
    timings = {}
    try:
        import cpu_ref
        x = np.full(c["nex"] * c["ney"] * c["nez"], c["volfrac"])
        for n in cands:
            cr = cpu_ref.CpuRef(c["nex"] + 1, c["ney"] + 1, c["nez"] + 1, xc=c["xc"], nu=c["nu"], nthreads=n)
            try:
                cr.assemble(x, c["Emin"], c["Emax"], c["penal"])
                cr.mg_setup(nlv)
                _, _, _, sec = cr.fgmres_mg(rtol=1e-30, maxits=2)
                timings[n] = float(sec)
            finally:
                cr.close()
        best = min(timings, key=timings.get)
        return best, {"seconds_for_2_fgmres_iterations": {str(k): round(v, 3) for k, v in timings.items()}}
    except Exception as e:  # calibration is a convenience, never a reason to lose the arm
        return cores, {"calibration_failed": repr(e)[:200]}


def cpu_whole_steps(c, steps, warmup, budget_s):
    """Whole design iterations of the reference's algorithm on the host cores (oracle/cpu_design_loop.py; thread count = the
    fastest of cores, cores/2, cores/4 for this problem on this box, see pick_cpu_threads).  Runs `warmup`
    untimed and up to `steps` timed iterations, stopping early (after at least one) when the next one would not fit into
    `budget_s` seconds of wall clock.  Nothing is extrapolated."""
    if c["filter"] != 1 or c["bc"] != "cantilever":
        return {"unavailable": "the CPU restatement covers the reference's own problem (cantilever, cone filter) only"}
    # threads pinned to cores, memory first-touched by the threads that use it (the solver is memory-bound)
    os.environ.setdefault("OMP_PROC_BIND", "spread")  # (fewer threads than cores: spread them over the sockets' memory channels)
    os.environ.setdefault("OMP_PLACES", "cores")
    import cpu_design_loop as cd
    nlv = min(4, gmg_levels(c))  # -nlvls 4 is the reference default (TopOpt.cc:116)
    cores, calib = pick_cpu_threads(c, host_cores(), nlv)  # (outside the timed steps and outside their budget)
    t_begin = time.perf_counter()
    dx = (c["xc"][1] - c["xc"][0]) / c["nex"]
    L = cd.CpuDesignLoop(c["nex"], c["ney"], c["nez"], volfrac=c["volfrac"], penal=c["penal"], Emin=c["Emin"], Emax=c["Emax"],
                         rmin=c["rmin_cells"] * dx, nlvls=nlv, rtol=c["rtol"], Xmin=c["Xmin"], Xmax=c["Xmax"],
                         movlim=c["movlim"], nthreads=cores, xc=c["xc"], nu=c["nu"])
    t_setup = time.perf_counter() - t_begin
    recs, last = [], 0.0
    for s in range(warmup + steps):
        if s > warmup and (time.perf_counter() - t_begin) + 1.1 * last > budget_s:
            break
        r = L.step()
        last = r["seconds"]["total"]
        if s >= warmup:
            recs.append(r)
    L.close()
    tot = sum(r["seconds"]["total"] for r in recs)
    parts = {k: float(np.mean([r["seconds"][k] for r in recs])) for k in recs[0]["seconds"]}
    return {"value": len(recs) / tot, "unit": "design iterations/s", "cores": cores, "host_cores": host_cores(), "thread_calibration": calib,
            "kind": "port",
            "sample": f"{len(recs)} WHOLE design iteration(s) (design iterations {recs[0]['itr']}..{recs[-1]['itr']} from the uniform "
                      f"start) of {c['desc'].split(',')[0]}: CSR re-assembly + Galerkin PtAP set-up ({nlv} levels) + FGMRES(100) "
                      "rtol 1e-5 warm-started with a multiplicative V-cycle, GMRES(4)+SOR smoothers, GMRES(30)+SOR coarse solve "
                      f"(LinearElasticity.cc:613-746 restated in oracle/c/cpu_ref.c, OpenMP, one thread = one PETSc rank) + cone "
                      f"filters + {L.mma_kind}; no extrapolation",
            "steps_timed": len(recs), "seconds_per_step": tot / len(recs), "seconds": parts,
            "krylov_iterations_per_step": [r["krylov_its"] for r in recs], "setup_seconds": t_setup,
            "fx": [r["fx"] for r in recs]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    c = dict(BASE, **WORKLOADS[args.workload])
    budget = float(os.environ.get("TOB_REF_BUDGET_S", "150"))
    cb = cpu_whole_steps(c, steps=max(1, args.steps), warmup=min(args.warmup, 1), budget_s=budget)
    if "unavailable" in cb:
        print(json.dumps({"impl": "reference", "unavailable": cb["unavailable"]}))
        return
    n = cb["steps_timed"]
    out = {"impl": "reference", "metric": "design_iters_per_sec", "value": cb["value"], "unit": "design iterations/s (whole job)",
           "n_gpus": args.gpus, "steps": n, "steps_requested": args.steps, "warmup": min(args.warmup, 1),
           "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": c["desc"], "elements": [c["nex"], c["ney"], c["nez"]],
                      "note": "PETSc/MPI are not installable here; this arm runs WHOLE design iterations of the reference's "
                              "algorithm restated in C/OpenMP (FGMRES + Galerkin PCMG + GMRES/SOR smoothers on the assembled "
                              "CSR matrix, oracle/c/cpu_ref.c) on the host cores.  It times as many of the requested steps as "
                              f"fit into {budget:.0f} s; `steps` is the number actually timed"},
           "cpu_baseline": cb, "gpu_launches": 0,
           "e2e": {"value": cb["value"], "unit": "design iterations/s (whole job)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out["wall_seconds"] = time.time() - t0
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gmg", action="store_true", help="skip the extra multigrid-solver leg")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the configs[2] (384x192x192 multigrid) leg")
    ap.add_argument("--no-spmv256", action="store_true", help="skip the 256^3 SpMV leg")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 = BASELINE configs[1] (default, the headline); cfg3/cfg4/cfg5 = configs[2..4]")
    ap.add_argument("--solver", default="auto", choices=["auto", "jacobi", "gmg"], help="state-solve preconditioner of the main arm")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3  # timing rule: at least 3 warm-up steps
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
