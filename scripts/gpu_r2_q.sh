#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sr_probe.py 2>&1 | grep -e "---" -e "cg trace" -e "its" | tee gpurun_out/r02_sr_probe.txt
timeout 300 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "sr or single_reduction or tf32 or galerkin" 2>&1 | tail -2
for tf in 0 1; do TOB_GMG_TF32=$tf TOB_GMG_TIMING=1 timeout 600 python scripts/gmg_growth_probe.py cfg3 8 2:3 2>&1 | grep -e "V-cycle" -e workload | tail -2 | cut -c1-600; done | tee gpurun_out/r02_tf32_smoother_gen3d.txt
