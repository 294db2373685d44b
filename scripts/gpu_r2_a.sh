#!/bin/bash
# round 2, run A: parity of the gen-2 z-march kernels + SpMV probe (cfg2 and 256^3)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "spmv_parity or linearity" 2>&1 | tail -5
timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 7,10,11 --reps 30 2>&1 | tee gpurun_out/r02_spmv_probe_a.txt | cut -c1-900
