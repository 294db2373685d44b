"""GPU probe: FP64 pipe micro-benchmarks + matrix-free SpMV GB/s for each kernel variant.
Usage: python scripts/spmv_probe.py [--sizes 192x96x96,256x256x256] [--reps 20] [--kernels 1,2,3]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from topopt_in_petsc_b200 import Context, LinearElasticity, solver_opts

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="192x96x96,256x256x256")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--kernels", default="1,2,3,4,5,6,7,8")
ap.add_argument("--micro", type=int, default=1)
ap.add_argument("--dens", default="uniform")
a = ap.parse_args()
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
out = {"hbm_peak_gbs": peaks["hbm_gbs"]}
if a.micro:
    ctx = Context(9, 5, 5)
    out["fp64_dfma_tflops"] = ctx.microbench_fp64(0)
    out["fp64_dmma_tflops"] = ctx.microbench_fp64(1)
    out["fp64_both_tflops"] = ctx.microbench_fp64(2)
    out["copy_gbs_1GiB"] = ctx.microbench_copy(1 << 30)
    ctx.close()
    print(json.dumps(out), flush=True)
names = {1: "naive_dense", 2: "zmarch_ty8", 3: "zmarch_ty16", 4: "zmarch_ty12_regs", 5: "zmarch_ty8_regs", 6: "zmarch_ty8_mbar", 7: "zmarch_ty16_mbar", 8: "zmarch_ty12_mbar", 9: "zmarch_ty16_regs", 10: "zm4_ty16_unroll4", 11: "zm4_ty16_merged_producer"}
for sz in a.sizes.split(","):
    ex, ey, ez = [int(v) for v in sz.split("x")]
    ctx = Context(ex + 1, ey + 1, ez + 1)
    phys = LinearElasticity(ctx)
    rng = np.random.default_rng(4321)
    x = np.full(ctx.nel, 0.12) if a.dens == "uniform" else rng.uniform(0, 1, ctx.nel)
    phys.SetDensity(x, 1e-9, 1.0, 3.0)
    u = np.random.default_rng(1234).uniform(-1, 1, ctx.ndof)
    y = phys.SpMV(u)  # leaves u resident in the operator's input buffer
    alg_bytes = 8 * ctx.nel + 16 * ctx.ndof
    rec = {"size": sz, "nel": ctx.nel, "ndof": ctx.ndof, "alg_bytes": alg_bytes, "kernels": {}}
    for k in [int(v) for v in a.kernels.split(",")]:
        reps = a.reps if k != 1 else max(2, a.reps // 5)
        ms = phys.SpMVBench(kernel=k, reps=reps)
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        rec["kernels"][names[k]] = {"ms": ms, "gbs": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"]}
    if os.environ.get("TOB_PROBE_TF32"):
        phys.SpMVtf32(u)
        ms = phys.SpMVtf32Bench(reps=a.reps)
        b32 = 8 * ctx.nel + 8 * ctx.ndof  # fp32 vectors, fp64 modulus as stored
        rec["kernels"]["tc32_tcgen05_tf32"] = {"ms": ms, "gbs_fp32_bytes": b32 / (ms * 1e-3) / 1e9, "alg_bytes_fp32": b32,
                                               "frac_of_measured_hbm": b32 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    out.setdefault("spmv", []).append(rec)
    print(json.dumps(rec), flush=True)
    phys.close(); ctx.close()
json.dump(out, open("gpurun_out/spmv_probe.json", "w"), indent=1)
