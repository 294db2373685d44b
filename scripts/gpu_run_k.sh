#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "spmv_parity" 2>&1 | tail -3
timeout 600 python scripts/spmv_probe.py --micro 0 --kernels 4,5,9 2>&1 | tee gpurun_out/k_probe.txt | cut -c1-700
