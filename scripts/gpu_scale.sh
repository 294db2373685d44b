#!/bin/bash
# usage: gpu_scale.sh N   -- default bench line at N ranks (+ the classical-PCG main arm for comparison)
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -2 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
print('N=$N',{k:d[k] for k in ('value','ms_per_step','gpu_launches')},'e2e',d['e2e']['value'],'spmv_us',d['roofline']['avg_launch_us'],'its',d['config']['pcg_iterations_per_step'])
print('gmg',d['gmg_solver'].get('value'),d['gmg_solver'].get('pcg_iterations_per_step'),d['gmg_solver'].get('error'))
print('cfg3',d['cfg3_gmg'])
PY
TOB_PCG_SR=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 8 --warmup 3 --no-gmg --no-cfg3 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('classical PCG N=$N:', {k:d[k] for k in ('value','ms_per_step')})"
