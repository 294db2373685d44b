#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "gmg" -s 2>&1 | tail -12
timeout 900 python scripts/gmg_probe.py 96x48x48 2>&1 | tee gpurun_out/e_gmg_probe.txt | tail -40
