#!/bin/bash
mkdir -p gpurun_out
echo "== edge-case tests"
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "tiny or error" 2>&1 | tail -5
echo "== ncu launch list of the bench command (3000 launches inside the timed solve)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 3000 --csv --log-file gpurun_out/l_bench_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-gmg > gpurun_out/l_ncu_bench.log 2>&1; tail -2 gpurun_out/l_ncu_bench.log | cut -c1-200
echo "== ncu full capture of the SpMV inside the solver (cfg2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zmarch -s 200 -c 1 -o gpurun_out/l_zmarch_cfg2 \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-gmg > gpurun_out/l_ncu_full.log 2>&1; tail -2 gpurun_out/l_ncu_full.log | cut -c1-200
echo "== bench N=1"
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/l_bench1_err.txt | grep "^{" | tee gpurun_out/l_bench1.json | cut -c1-300
tail -3 gpurun_out/l_bench1_err.txt
