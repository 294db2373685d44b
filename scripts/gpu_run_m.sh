#!/bin/bash
mkdir -p gpurun_out
echo "== solver tests (fused PCG path)"
timeout 900 python -m pytest tests/test_hex8_gpu.py tests/test_filter_mma_gpu.py -m gpu -x -q -k "golden or solve or custom or design_loop or tiny" 2>&1 | tail -4
echo "== bench N=1"
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu 2>gpurun_out/m_bench1_err.txt | grep "^{" | tee gpurun_out/m_bench1.json | cut -c1-200
tail -3 gpurun_out/m_bench1_err.txt
