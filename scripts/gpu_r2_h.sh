#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "tf32" 2>&1 | tail -6
TOB_PROBE_TF32=1 TOB_TC32_GEN=2 timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 2>&1 | tee gpurun_out/r02_spmv_probe_h.txt | cut -c1-700
