#!/bin/bash
# final multi-GPU lines: full bench line at N ranks (8 timed steps), plus the one-rank main arm with the same step count on the same box
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -2 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
print('N=$N',{k:d[k] for k in ('value','ms_per_step','gpu_launches')},'e2e',d['e2e']['value'],'spmv_us',d['roofline']['avg_launch_us'],'its',d['config']['pcg_iterations_per_step'])
print('gmg',d['gmg_solver'].get('value'),d['gmg_solver'].get('pcg_iterations_per_step'),d['gmg_solver'].get('error'))
print('cfg3',d['cfg3_gmg'].get('value'),d['cfg3_gmg'].get('pcg_iterations_per_step'))
PY
if [ "$2" = "one" ]; then
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-spmv256 2>/dev/null | grep "^{" > gpurun_out/r02_bench_n1_8steps.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1_8steps.json').read()); print('N=1 8 steps:', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'gmg', d['gmg_solver']['value'], 'cfg3', d['cfg3_gmg']['value'])"
fi
