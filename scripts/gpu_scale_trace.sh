#!/bin/bash
# usage: gpu_scale_trace.sh N -- time stamps of the SR-PCG phases at N ranks, then the default bench line
N=$1
mkdir -p gpurun_out
TOB_CG_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2> gpurun_out/r02_cg_trace_n$N.err | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N traced:', {k:d[k] for k in ('value','ms_per_step')}, d['config']['pcg_iterations_per_step'])"
grep "cg trace" gpurun_out/r02_cg_trace_n$N.err | head -$N | tee gpurun_out/r02_cg_trace_n$N.txt
bash scripts/gpu_scale.sh $N
