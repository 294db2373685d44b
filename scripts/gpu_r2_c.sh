#!/bin/bash
# round 2, run C: new parity tests (Galerkin levels, single-reduction PCG, cfg2-size, trajectory, restart) + default bench line
mkdir -p gpurun_out
nproc; free -g | head -2 | tail -1
timeout 900 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "galerkin or single_reduction or split_objective or gmg_pcg" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -x -q -k "trajectory or reference_written" 2>&1 | tail -8
timeout 1200 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -s -k "cfg2" 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -3 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_c.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e',d['e2e']['value']); print('roofline',{k:d['roofline'][k] for k in ('frac','avg_launch_us','share_of_step')})
print('spmv256',d['spmv_256']); print('gmg',d['gmg_solver']); print('cfg3',d['cfg3_gmg'])
print('cpu',{k:d['cpu_baseline'].get(k) for k in ('value','cores','seconds_per_step','krylov_iterations_per_step','seconds')})
print('its',d['config']['pcg_iterations_per_step'])
PY
