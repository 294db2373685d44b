#!/bin/bash
# full GPU suite + smoke + tensor-core kernel timing (gen 3 with look-ahead) + its ncu
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
rm -f gpurun_out/r02_spmv_probe_m.txt
TOB_PROBE_TF32=1 TOB_TC32_GEN=3 timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 2>&1 | cut -c1-700 | tee -a gpurun_out/r02_spmv_probe_m.txt
for tf in 0 1; do
  TOB_GMG_TF32=$tf TOB_GMG_TIMING=1 timeout 600 python scripts/gmg_growth_probe.py cfg3 8 2:3 2>&1 | grep "gmg timing\|^{\|rror" | cut -c1-600 | sed "s/^/tf32=$tf /" | tee -a gpurun_out/r02_tf32_smoother_gen3b.txt
done
TOB_PROBE_TF32=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_tc32c -s 4 -c 1 -o gpurun_out/r02_tc32c_256 python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels 10 --micro 0 > gpurun_out/r02_ncu_tc32c.log 2>&1
tail -1 gpurun_out/r02_ncu_tc32c.log | cut -c1-200
