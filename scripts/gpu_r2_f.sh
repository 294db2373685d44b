#!/bin/bash
# round 2, run F: tcgen05 TF32 operator (parity + timing), dense coarse inverse, regression of the GMG tests, in-solver ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "tf32 or galerkin or gmg_pcg or tiny" 2>&1 | tail -6
TOB_PROBE_TF32=1 timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 2>&1 | tee gpurun_out/r02_spmv_probe_f.txt | cut -c1-700
TOB_GMG_TIMING=1 timeout 300 python scripts/gmg_growth_probe.py cfg2 20 2:3 2>&1 | grep "gmg timing\|^{" | cut -c1-700 | tee gpurun_out/r02_gmg_growth_cfg2_dense.txt
TOB_GMG_TIMING=1 timeout 300 python scripts/gmg_growth_probe.py cfg3 12 2:3 2>&1 | grep "gmg timing\|^{" | cut -c1-700 | tee gpurun_out/r02_gmg_growth_cfg3_dense.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zm4 -s 3000 -c 1 -o gpurun_out/r02_zm4_insolver_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 > gpurun_out/r02_ncu_insolver.log 2>&1
tail -2 gpurun_out/r02_ncu_insolver.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_bench_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 > gpurun_out/r02_ncu_launches.log 2>&1
tail -1 gpurun_out/r02_ncu_launches.log | cut -c1-200
