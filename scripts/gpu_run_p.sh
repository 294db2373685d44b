#!/bin/bash
# parity of all SpMV variants + timing A/B: block constants from the constant bank (default build) vs shared memory
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 2,3,5,9,6,7,8 --reps 20 2>&1 | tee gpurun_out/p_probe_const.txt | cut -c1-1200
echo "--- shared-memory constants build"
TOB_LIB_PATH=$PWD/build/libtopopt_b200_lds.so timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 5,9,6,7 --reps 20 2>&1 | tee gpurun_out/p_probe_lds.txt | cut -c1-900
