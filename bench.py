#!/usr/bin/env python
"""bench.py -- design iterations/sec (and matrix-free SpMV GB/s) of the B200-native hot path.

A "step" is one design iteration of main.cc:54-105 (state solve + objective/sensitivities, sensitivity
filtering, MMA update, density filter, MND) on the synthetic cantilever of BASELINE.json configs[1]:
192x96x96 elements, matrix-free PCG + Jacobi (rtol 1e-5, warm start), density filter rmin = 3 cells.

  python bench.py --gpus 1 --steps K --warmup W              (our arm; N>1 under torchrun, z-slab + NCCL)
  python bench.py --impl reference --steps K --warmup W      (reference arm: CPU restatement, rank 0 only)

`value` : device-resident design iterations/sec (inputs already in HBM), CUDA events, max over ranks.
`e2e`   : the same loop driven through the C-ABI with HOST (pinned) buffers -- every call stages its
          vectors host<->device inside the timed region, exactly what the reference's main.cc/MMA.cc
          do through VecGetArray.
`roofline`: the fine-level matrix-free SpMV kernel, timed live with CUDA event pairs around every 8th launch of the
          eager part of the timed region (pairs around every launch cost 6 % of a PCG iteration); algorithmic
          bytes = 8*nel + 16*ndof per application.
`gmg_solver`: the same design iterations with the multigrid-preconditioned solver (secondary leg, N >= 1).
`cpu_baseline`: the C/OpenMP restatement of the reference's assembled-CSR path (oracle/c/cpu_ref.c) timed
          on the host cores on a bounded sample (see `sample`).
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

CFG = dict(nex=192, ney=96, nez=96, volfrac=0.12, penal=3.0, Emin=1e-9, Emax=1.0, rmin_cells=3.0, Xmin=0.0, Xmax=1.0,
           movlim=0.2, rtol=1e-5, nu=0.3)
WORKLOAD = "cantilever 192x96x96, matrix-free PCG+Jacobi rtol 1e-5 warm-started, density filter rmin=3 cells, MMA m=1"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


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


class DesignLoop:
    """main.cc:48-105 over the C-ABI.  resident=True keeps every vector in HBM (DevVec); resident=False uses
    pinned host numpy buffers, so each C-ABI call stages host<->device inside the timed region."""

    def __init__(self, ctx, resident, profile=True, pc=0, nlvls=4):
        from topopt_in_petsc_b200 import Filter, LinearElasticity, MMA, solver_opts
        self.ctx, self.resident = ctx, resident
        c = CFG
        self.phys = LinearElasticity(ctx, nu=c["nu"], opts=solver_opts(pc=pc, nlvls=nlvls, rtol=c["rtol"], cheb_degree=3,
                                                                       coarse_its=20 if nlvls >= 5 else 60,
                                                                       maxits=100000 if pc == 0 else 500, check_every=50))
        # every 8th fine-level SpMV launch is bracketed by CUDA events (timing all of them costs ~6 % of the step)
        self.phys.SetProfile(0 if (not profile or os.environ.get("TOB_BENCH_NOPROFILE")) else 8)
        dx = 2.0 / c["nex"]
        self.flt = Filter(ctx, 1, c["rmin_cells"] * dx)
        self.mma = MMA(ctx, 1)
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
        self.flt.FilterProject(self.v["x"], self.v["xTilde"], self.v["xPhys"], False, 0.1, 0.0)  # main.cc:48
        self.hist = []

    def step(self):
        c, v, n8 = CFG, self.v, 8 * self.n
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
        self.flt.Gradients(v["x"], v["xTilde"], v["dfdx"], 1, [v["dgdx"]], False, 0.1, 0.0)  # :76
        self.mma.SetOuterMovelimit(c["Xmin"], c["Xmax"], c["movlim"], v["x"], v["xmin"], v["xmax"])  # :81
        self.mma.Update(v["x"], v["dfdx"], np.array([gx]), [v["dgdx"]], v["xmin"], v["xmax"])  # :85
        ch = self.mma.DesignChange(v["x"], v["xold"])  # :89
        self.flt.FilterProject(v["x"], v["xTilde"], v["xPhys"], False, 0.1, 0.0)  # :98
        mnd = self.flt.GetMND(v["xPhys"])  # :102
        if not self.resident:
            # bytes staged by the calls above (host pointers): inputs H2D / outputs D2H, counted per vector
            self.h2d += n8 * (1 + 1 + 4 + 1 + 5 + 2 + 1 + 1)
            self.d2h += n8 * (2 + 1 + 2 + 2 + 1 + 1 + 2)
        self.hist.append(dict(itr=self.itr, fx=fx, gx=gx, ch=ch, mnd=mnd, pcg_its=st.iterations))
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
    from topopt_in_petsc_b200 import Context, nccl_unique_id
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
    c = CFG
    stream = torch.cuda.current_stream().cuda_stream
    ctx = Context(c["nex"] + 1, c["ney"] + 1, c["nez"] + 1, rank=rank, nranks=world, device=local, nccl_id=nccl_id,
                  stream=stream if stream else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(loop, K, W, sampler=None):
        for _ in range(W):
            loop.step()
        loop.spmv_s, loop.spmv_n, loop.spmv_calls = 0.0, 0, 0
        loop.h2d = loop.d2h = 0
        nwarm = len(loop.pcg_its)
        ctx.launch_count(reset=True)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            loop.step()
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks, loop.pcg_its[nwarm:]

    K, W = args.steps, args.warmup
    # deepest hierarchy (<= 6 levels) whose coarsening divides the grid; on N > 1 ranks the levels whose slab
    # boundaries are no longer aligned (or that are small) are replicated on every rank, so the depth does not depend on N
    # (profiles/r01_gmg_sweep_192x96x96.txt: 6 levels + 20 coarse iterations beat 4 levels + 60 by 2-3x)
    nlv = 6
    while nlv > 2 and any(d % (1 << (nlv - 1)) for d in (c["nex"], c["ney"], c["nez"])):
        nlv -= 1
    main_pc = 1 if args.solver == "gmg" else 0
    # ---- device-resident arm -----------------------------------------------------------------
    loop = DesignLoop(ctx, resident=True, pc=main_pc, nlvls=nlv)
    ms, clocks, its = timed(loop, K, W, ClockSampler(local))
    launches = ctx.launch_count()
    spmv_s, spmv_n, spmv_calls = loop.spmv_s, loop.spmv_n, loop.spmv_calls
    hist = loop.hist
    loop.close()
    # ---- end-to-end arm: host (pinned) buffers through the same C-ABI calls ---------------------------
    loop2 = DesignLoop(ctx, resident=False, pc=main_pc, nlvls=nlv)
    ms2, _, _ = timed(loop2, K, W)
    h2d, d2h = loop2.h2d // K, loop2.d2h // K
    loop2.close()

    # ---- same job with the multigrid-preconditioned solver (not the BASELINE-named cfg2 solver; reported beside it) ----
    gmg = None
    if not args.no_gmg and main_pc == 0:
        try:  # a secondary leg must never take the headline measurement down with it
            loop3 = DesignLoop(ctx, resident=True, pc=1, nlvls=nlv)
            ms3, _, its3 = timed(loop3, K, W)
            gmg = {"value": K / (ms3 * 1e-3), "unit": "design iterations/s (whole job)", "ms_per_step": ms3 / K, "nlvls": nlv,
                   "pcg_iterations_per_step": its3, "smoother": "Chebyshev-Jacobi degree 3",
                   "coarse": f"{20 if nlv >= 5 else 60} Jacobi-PCG iterations",
                   "final_fx": loop3.hist[-1]["fx"]}
            loop3.close()
        except Exception as e:  # noqa: BLE001
            gmg = {"error": str(e)[:300]}

    if rank == 0:
        peak, kind = peaks()
        nel, ndof = ctx.nel, ctx.ndof
        # per-launch algorithmic bytes of THIS rank's slab (SpMV is launched once per rank per application)
        nel_loc, ndof_loc = ctx.nel_owned, ctx.ndof_owned
        alg = 8 * nel_loc + 16 * ndof_loc
        achieved = alg * spmv_n / spmv_s / 1e9 if spmv_s > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
        if os.path.exists(tp) and world == 1:
            traffic = json.load(open(tp)).get("192x96x96", {}).get("dram_bytes_per_launch")
        out = {
            "metric": "design_iters_per_sec", "value": K / (ms * 1e-3), "unit": "design iterations/s (whole job)",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "elements": [c["nex"], c["ney"], c["nez"]], "nel": nel, "ndof": ndof,
                       "partition": f"z-slab x{world}", "l2": "solver working set (6 vectors x 44 MB + modulus) exceeds the 126 MB L2; "
                       "every SpMV is separated by >300 MB of vector-update traffic",
                       "pcg_iterations_per_step": its, "timed_design_iterations": [h["itr"] for h in hist[W:]]},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": K / (ms2 * 1e-3), "unit": "design iterations/s (whole job)", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms2 / K},
            "roofline": {"kernel": "k_hex8_zmarch (fine-level matrix-free hex8 SpMV, fp64)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_kind": kind, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg, "launches_timed": spmv_n, "launches_total": spmv_calls,
                         "avg_launch_us": (spmv_s / spmv_n * 1e6) if spmv_n else None,
                         "share_of_step": (spmv_s / spmv_n * spmv_calls / (ms * 1e-3)) if spmv_n else None},
            "history": [{k: (round(v, 9) if isinstance(v, float) else v) for k, v in h.items()} for h in hist[W:]],
            "gmg_solver": gmg,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_sample(steps=1, warmup=0, its_full=int(np.mean(its)) if its else None)
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_sample(steps, warmup, its_full=None, sample_its=10):
    """Reference CPU path on a bounded sample: per step one full CSR assembly (the reference re-assembles every
    design iteration, LinearElasticity.cc:190), `sample_its` Jacobi-PCG iterations on the assembled matrix and one
    objective/sensitivity pass, at the full 192x96x96 size.  Iterations/sec is extrapolated with the number of
    Krylov iterations the identical algorithm needs per design iteration (measured on the GPU arm, or the
    committed profiles/cfg2_pcg_iterations.json when the reference arm runs alone)."""
    import cpu_ref
    c = CFG
    # all physical host cores, threads pinned (the CSR SpMV is memory-bound: first-touch placement matters)
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    os.environ.setdefault("OMP_PLACES", "cores")
    try:
        import psutil
        cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        cores = os.cpu_count() or 1
    if its_full is None:
        p = os.path.join(ROOT, "profiles", "cfg2_pcg_iterations.json")
        its_full = json.load(open(p))["mean_pcg_iterations_per_design_iteration"] if os.path.exists(p) else 1500
    cr = cpu_ref.CpuRef(c["nex"] + 1, c["ney"] + 1, c["nez"] + 1, nu=c["nu"], nthreads=cores)
    x = np.full(cr.nel, c["volfrac"])
    times = []
    for s in range(warmup + steps):
        t_asm = cr.assemble(x, c["Emin"], c["Emax"], c["penal"])
        U, its, rr, t_pcg = cr.pcg(rtol=0.0, maxit=sample_its)
        _, _, _, t_obj = cr.objective(U, x, c["Emin"], c["Emax"], c["penal"], c["volfrac"])
        u = np.ones(cr.ndof)
        _, t_spmv = cr.spmv(u, reps=3)
        if s >= warmup:
            times.append((t_asm, t_pcg / max(its, 1), t_obj, t_spmv))
    cr.close()
    t_asm, t_it, t_obj, t_spmv = [float(np.mean([t[k] for t in times])) for k in range(4)]
    t_full = t_asm + its_full * t_it + t_obj
    nel, ndof = c["nex"] * c["ney"] * c["nez"], 3 * (c["nex"] + 1) * (c["ney"] + 1) * (c["nez"] + 1)
    return {"value": 1.0 / t_full, "unit": "design iterations/s", "cores": cores, "kind": "port",
            "sample": f"192x96x96: 1 CSR assembly + {sample_its} Jacobi-PCG iterations + 1 objective pass per step, extrapolated to "
                      f"{its_full} PCG iterations per design iteration (same algorithm and tolerance as the GPU arm); filter and "
                      "MMA excluded (favours the CPU arm)",
            "seconds": {"assembly": t_asm, "pcg_iteration": t_it, "objective": t_obj, "csr_spmv": t_spmv},
            "csr_spmv_gbs_algorithmic": (8 * nel + 16 * ndof) / t_spmv / 1e9,
            "sample_seconds_per_step": t_asm + sample_its * t_it + t_obj}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    cb = cpu_sample(steps=max(1, args.steps), warmup=min(args.warmup, 1))
    c = CFG
    out = {"impl": "reference", "metric": "design_iters_per_sec", "value": cb["value"], "unit": "design iterations/s (whole job)",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"],
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "elements": [c["nex"], c["ney"], c["nez"]],
                      "note": "PETSc/MPI are not installable here; this arm is the C/OpenMP restatement of the reference's "
                              "assembled-CSR path (oracle/c/cpu_ref.c) on the host cores.  The reference's own sources do run "
                              "against oracle/petsc_lite (oracle/_ref/topopt_ref), but serially with std::map matrices: a "
                              "fidelity tool that pins the oracles, not a timing baseline"},
           "cpu_baseline": cb, "gpu_launches": 0, "wall_seconds": None,
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
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"],
                    help="cfg2 = BASELINE configs[1] 192x96x96 (default, the headline); cfg3 = configs[2] 384x192x192 (scaling studies)")
    ap.add_argument("--solver", default="jacobi", choices=["jacobi", "gmg"], help="state-solve preconditioner of the main arm")
    a = ap.parse_args()
    os.environ["NCCL_DEBUG"] = os.environ.get("TOB_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    if a.workload == "cfg3":
        CFG.update(nex=384, ney=192, nez=192)
        WORKLOAD = "cantilever 384x192x192, matrix-free PCG + geometric multigrid V-cycle, density filter rmin=3 cells, MMA m=1"
        a.solver = "gmg"
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3  # timing rule: at least 3 warm-up steps
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
