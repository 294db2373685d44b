#!/bin/bash
# usage: gpu_trace_only.sh N [tests]  -- SR-PCG phase time stamps at N ranks (3 design iterations of cfg2); optional 2-rank parity tests first
N=$1
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_dropin_gpu.py -m gpu -x -q -k "two_rank" 2>&1 | tail -3; fi
TOB_CG_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2> gpurun_out/r02_cg_trace_n$N.err | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N traced:', {k:d[k] for k in ('value','ms_per_step')}, d['config']['pcg_iterations_per_step'])"
grep "cg trace" gpurun_out/r02_cg_trace_n$N.err | head -$N | tee gpurun_out/r02_cg_trace_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N untraced 8 steps:', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])"
