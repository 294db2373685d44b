#!/bin/bash
# 2 GPUs: programmatic dependent launch on/off (TOB_PDL), parity first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_dropin_gpu.py -m gpu -x -q -k "two_rank" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "single_reduction" 2>&1 | tail -2
for pdl in 1 0; do
TOB_PDL=$pdl TOB_CG_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$pdl bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2> gpurun_out/r02_cg_trace_pdl$pdl.err | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('PDL=$pdl N=2 traced:', {k:d[k] for k in ('value','ms_per_step')}, d['config']['pcg_iterations_per_step'])"
grep "cg trace" gpurun_out/r02_cg_trace_pdl$pdl.err | head -2
tail -3 gpurun_out/r02_cg_trace_pdl$pdl.err | cut -c1-300
done
