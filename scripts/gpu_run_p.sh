#!/bin/bash
# parity of all SpMV variants + timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 5,9,6,7,8 --reps 20 2>&1 | tee gpurun_out/p_probe.txt | cut -c1-1200
for k in ${NCU_KERNELS:-7}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zmarch -s 4 -c 1 -o gpurun_out/p_zmarch_k$k \
   python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels $k --micro 0 > gpurun_out/p_ncu_k$k.log 2>&1
done
