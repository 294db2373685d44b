#!/bin/bash
# 2 GPUs: device time stamps of the single-reduction PCG (TOB_CG_TRACE), then the N=2 bench line
mkdir -p gpurun_out
TOB_CG_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2> gpurun_out/r02_cg_trace_n2.err | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 traced:', {k:d[k] for k in ('value','ms_per_step')}, d['config']['pcg_iterations_per_step'])"
grep "cg trace" gpurun_out/r02_cg_trace_n2.err | tail -4 | tee gpurun_out/r02_cg_trace_n2.txt
TOB_CG_TRACE=1 TOB_PCG_SR=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2>&1 | grep "cg trace" | tail -2 | tee gpurun_out/r02_cg_trace_n1.txt
bash scripts/gpu_scale.sh 2
