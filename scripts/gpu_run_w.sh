#!/bin/bash
# tiled cone filter: parity, timing against the direct-gather kernel, ncu of both new default kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_filter_mma_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python scripts/filter_probe.py 2>&1 | tail -1 | tee gpurun_out/w_filter_tiled.json
TOB_CONE_TILED=0 timeout 120 python scripts/filter_probe.py 2>&1 | tail -1 | tee gpurun_out/w_filter_gather.json
timeout 120 python scripts/filter_probe.py 48 24 24 1 2>&1 | tail -1
TOB_CONE_TILED=0 timeout 120 python scripts/filter_probe.py 48 24 24 1 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cone -s 2 -c 1 -o gpurun_out/w_cone_tiled python scripts/filter_probe.py > gpurun_out/w_ncu_cone.log 2>&1
TOB_CONE_TILED=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_cone -s 2 -c 1 python scripts/filter_probe.py 2>&1 | grep -E "k_cone|duration|dram__" | head -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zmarch -s 4 -c 1 -o gpurun_out/w_zmarch_k7 python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels 7 --micro 0 > gpurun_out/w_ncu_k7.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-gmg 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, 'fx', d['history'][-1]['fx'])"
