#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "galerkin or gmg_pcg or tiny" 2>&1 | tail -4
timeout 300 python scripts/gmg_growth_probe.py cfg2 8 2:3 2>&1 | grep "^{\|rror" | cut -c1-400
timeout 300 python scripts/gmg_growth_probe.py cfg3 8 2:3 2>&1 | grep "^{\|rror" | cut -c1-400
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg4',{k:d[k] for k in ('value','ms_per_step')}, d['config']['pcg_iterations_per_step'])"
