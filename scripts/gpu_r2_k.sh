#!/bin/bash
# gen-3 tensor-core kernel: parity + timing + smoother; fp32-dinv PCG regression; quick bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "tf32 or iteration1_compliance_golden or solve_obj_sens or single_reduction" 2>&1 | tail -5
for gen in 2 3; do TOB_PROBE_TF32=1 TOB_TC32_GEN=$gen timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 2>&1 | cut -c1-700 | sed "s/^/gen=$gen /" | tee -a gpurun_out/r02_spmv_probe_k.txt; done
rm -f gpurun_out/r02_tf32_smoother_gen3.txt
for tf in 0 1; do
  TOB_GMG_TF32=$tf TOB_GMG_TIMING=1 timeout 600 python scripts/gmg_growth_probe.py cfg3 8 2:3 2>&1 | grep "gmg timing\|^{\|rror" | cut -c1-600 | sed "s/^/tf32=$tf /" | tee -a gpurun_out/r02_tf32_smoother_gen3.txt
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-cfg3 --no-spmv256 2>/dev/null | grep "^{" > gpurun_out/r02_bench_k.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_k.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['roofline']['avg_launch_us'], d['config']['pcg_iterations_per_step'], 'gmg', d['gmg_solver']['value'])
PY
