#!/bin/bash
# round 2, run B: gen-2 kernel with block sparsity + edge-row skipping: parity, probe, ncu (full set + source)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "spmv_parity or linearity" 2>&1 | tail -3
timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 7,10,11 --reps 30 2>&1 | tee gpurun_out/r02_spmv_probe_b.txt | cut -c1-900
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zm4 -s 4 -c 1 -o gpurun_out/r02_zm4_k10_256 python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels 10 --micro 0 > gpurun_out/r02_ncu_k10.log 2>&1
tail -2 gpurun_out/r02_ncu_k10.log
