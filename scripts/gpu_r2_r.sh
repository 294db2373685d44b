#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sr_probe.py 2>&1 | grep -e "---" -e "cg trace" -e "its" | tee gpurun_out/r02_sr_probe.txt
timeout 900 python -m pytest tests/test_hex8_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-spmv256 2>/dev/null | grep "^{" > gpurun_out/r02_bench_r.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_r.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')},'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['avg_launch_us'],'its',d['config']['pcg_iterations_per_step'])
print('gmg',d['gmg_solver']['value'],d['gmg_solver']['pcg_iterations_per_step'],'cfg3',d['cfg3_gmg'].get('value'),d['cfg3_gmg'].get('pcg_iterations_per_step'))
PY
