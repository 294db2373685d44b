#!/bin/bash
# TF32 tensor-core smoother inside the fp64 outer PCG: outer iteration counts with / without, cfg2 and cfg3
mkdir -p gpurun_out
for tf in 0 1; do
  TOB_GMG_TF32=$tf TOB_GMG_TIMING=1 timeout 600 python scripts/gmg_growth_probe.py cfg2 12 2:3 2>&1 | grep "gmg timing\|^{\|rror" | cut -c1-600 | sed "s/^/tf32=$tf /" | tee -a gpurun_out/r02_tf32_smoother.txt
  TOB_GMG_TF32=$tf TOB_GMG_TIMING=1 timeout 600 python scripts/gmg_growth_probe.py cfg3 8 2:3 2>&1 | grep "gmg timing\|^{\|rror" | cut -c1-600 | sed "s/^/tf32=$tf /" | tee -a gpurun_out/r02_tf32_smoother.txt
done
