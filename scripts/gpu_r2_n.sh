#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "tf32" 2>&1 | tail -3
TOB_PROBE_TF32=1 TOB_TC32_GEN=3 timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 2>&1 | cut -c1-700 | tee gpurun_out/r02_spmv_probe_n.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')},'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['avg_launch_us'],'its',d['config']['pcg_iterations_per_step'])
print('spmv256',d['spmv_256']['frac'],'gmg',d['gmg_solver']['value'],d['gmg_solver']['pcg_iterations_per_step'],'cfg3',d['cfg3_gmg'].get('value'),d['cfg3_gmg'].get('pcg_iterations_per_step'))
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['seconds_per_step'])
PY
TOB_PCG_LAZYX=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('lazyx off:', {k:d[k] for k in ('value','ms_per_step')})"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_bench_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 > gpurun_out/r02_ncu_launches.log 2>&1
