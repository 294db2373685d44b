#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gmg_growth_probe.py cfg2 20 2>&1 | grep "^{" | tee gpurun_out/r02_gmg_growth_cfg2.txt
timeout 900 python -m pytest tests/test_filter_mma_gpu.py -m gpu -x -q -s -k "pde" 2>&1 | tail -6
