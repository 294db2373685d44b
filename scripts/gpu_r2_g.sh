#!/bin/bash
# round 2, run G (2 GPUs): BASELINE configs[3] (cfg4) on 1 GPU, configs[4] (cfg5) on 2 GPUs, cfg2 at N=2 with the fenced SR path
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg4 --steps 4 --warmup 3 --no-cpu > gpurun_out/r02_bench_cfg4_n1.json 2> gpurun_out/r02_bench_cfg4_n1.err; tail -2 gpurun_out/r02_bench_cfg4_n1.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_bench_cfg4_n1.json') if l.startswith('{')][-1])
    print('cfg4',{k:d[k] for k in ('value','ms_per_step')}, 'e2e',d['e2e']['value'], 'its',d['config']['pcg_iterations_per_step'], [ (h['fx'],h['beta'],h['mnd']) for h in d['history']])
except Exception as e: print('cfg4 failed',e)
PY
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload cfg5 --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg5_n2.json 2> gpurun_out/r02_bench_cfg5_n2.err; tail -3 gpurun_out/r02_bench_cfg5_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_bench_cfg5_n2.json') if l.startswith('{')][-1])
    print('cfg5',{k:d[k] for k in ('value','ms_per_step')}, 'e2e',d['e2e']['value'], 'its',d['config']['pcg_iterations_per_step'], [h['fx'] for h in d['history']], d['roofline'])
except Exception as e: print('cfg5 failed',e)
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; tail -2 gpurun_out/r02_bench_n2b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n2b.json') if l.startswith('{')][-1])
print('N=2',{k:d[k] for k in ('value','ms_per_step')}, d['roofline']['avg_launch_us'], 'gmg',d['gmg_solver'].get('value'),d['gmg_solver'].get('pcg_iterations_per_step'),'cfg3',d['cfg3_gmg'])
PY
