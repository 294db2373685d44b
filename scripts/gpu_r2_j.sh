#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "tf32" 2>&1 | tail -4
TOB_PROBE_TF32=1 timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 --sizes 256x256x256 2>&1 | cut -c1-700 | tee gpurun_out/r02_spmv_probe_j.txt
TOB_PROBE_TF32=1 TOB_TC32_PACKED=0 timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 --sizes 256x256x256 2>&1 | cut -c1-700
rm -f gpurun_out/r02_tf32_smoother.txt
for tf in 0 1; do
  TOB_GMG_TF32=$tf timeout 600 python scripts/gmg_growth_probe.py cfg2 12 2:3 2>&1 | grep "^{\|rror" | cut -c1-600 | sed "s/^/tf32=$tf /" | tee -a gpurun_out/r02_tf32_smoother.txt
  TOB_GMG_TF32=$tf timeout 600 python scripts/gmg_growth_probe.py cfg3 8 2:3 2>&1 | grep "^{\|rror" | cut -c1-600 | sed "s/^/tf32=$tf /" | tee -a gpurun_out/r02_tf32_smoother.txt
done
TOB_PROBE_TF32=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_tc32b -s 4 -c 1 -o gpurun_out/r02_tc32b_256 python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels 10 --micro 0 > gpurun_out/r02_ncu_tc32b.log 2>&1
tail -1 gpurun_out/r02_ncu_tc32b.log | cut -c1-200
