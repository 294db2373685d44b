#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench N=1"; timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/b_bench_err.txt | tee gpurun_out/b_bench.json | cut -c1-3000
tail -5 gpurun_out/b_bench_err.txt
