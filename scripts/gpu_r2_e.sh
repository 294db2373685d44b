#!/bin/bash
# round 2, run E (2 GPUs): 2-rank parity (incl. single-reduction PCG), V-cycle timing, bench at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s 2>&1 | tail -8
TOB_GMG_TIMING=1 timeout 300 python scripts/gmg_growth_probe.py cfg2 3 2:3 2>&1 | grep "gmg timing\|^{" | cut -c1-600
TOB_GMG_TIMING=1 timeout 300 python scripts/gmg_growth_probe.py cfg3 3 2:3 2>&1 | grep "gmg timing\|^{" | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -3 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}); print('e2e',d['e2e']['value'])
print('roofline',{k:d['roofline'][k] for k in ('frac','avg_launch_us','share_of_step')})
print('gmg',d['gmg_solver']); print('cfg3',d['cfg3_gmg']); print('its',d['config']['pcg_iterations_per_step'])
PY
TOB_PCG_SR=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-gmg --no-cfg3 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('classical PCG N=2:', {k:d[k] for k in ('value','ms_per_step')}, d['config']['pcg_iterations_per_step'])"
