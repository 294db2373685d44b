#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gmg_probe2.py 192x96x96 2>&1 | tee gpurun_out/j_gmg_sweep.txt | tail -30
