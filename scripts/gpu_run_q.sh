#!/bin/bash
mkdir -p gpurun_out
for k in ${NCU_KERNELS:-7}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zmarch -s 4 -c 1 -o gpurun_out/q_zmarch_k$k \
   python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels $k --micro 0 > gpurun_out/q_ncu_k$k.log 2>&1
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/q_bench_err.txt | grep "^{" | tee gpurun_out/q_bench_cfg2_n1.json | cut -c1-1500
tail -3 gpurun_out/q_bench_err.txt
