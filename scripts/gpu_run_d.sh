#!/bin/bash
mkdir -p gpurun_out
echo "== quick correctness of flag kernels"
for k in 6 7 8; do timeout 120 python scripts/dbg_zm.py $k 2>&1 | tail -1; done
echo "== spmv parity tests"
timeout 900 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "spmv" 2>&1 | tail -5
echo "== probe (L2 flushed between reps)"
timeout 600 python scripts/spmv_probe.py --micro 0 --kernels 2,4,5,6,7,8 2>&1 | tee gpurun_out/d_probe.txt | cut -c1-1200
for k in ${NCU_KERNELS:-6}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zmarch -s 4 -c 1 -o gpurun_out/d_zmarch_k$k \
   python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels $k --micro 0 > gpurun_out/d_ncu_k$k.log 2>&1
done
