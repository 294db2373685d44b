#!/bin/bash
# confirm the final state on one GPU: full suite, smoke, SpMV probe (previous build beside it), default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 5,9,6,7,8 --reps 20 2>&1 | tee gpurun_out/v_probe.txt | cut -c1-1200
echo "--- previous build (no running output pointer)"
TOB_LIB_PATH=$PWD/build/libtopopt_b200_prev.so timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 7 --reps 20 2>&1 | tee gpurun_out/v_probe_prev.txt | cut -c1-500
echo "== default bench"
timeout 900 python bench.py 2>gpurun_out/v_bench_err.txt | grep "^{" | tee gpurun_out/v_bench_default.json | cut -c1-300
tail -2 gpurun_out/v_bench_err.txt
