#!/bin/bash
# 2 GPUs: multi-rank parity, SR-PCG trace after the residual elimination, N=2 bench line; tensor-core operator probe (gen 3, Y3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_dropin_gpu.py -m gpu -x -q -k "two_rank" 2>&1 | tail -4
TOB_CG_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2> gpurun_out/r02_cg_trace_n2.err | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 traced:', {k:d[k] for k in ('value','ms_per_step')}, d['config']['pcg_iterations_per_step'])"
grep "cg trace" gpurun_out/r02_cg_trace_n2.err | head -4 | tee gpurun_out/r02_cg_trace_n2.txt
bash scripts/gpu_scale.sh 2
timeout 300 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "tf32" 2>&1 | tail -2
for y3 in 1 0; do TOB_TC32_Y3=$y3 TOB_PROBE_TF32=1 TOB_TC32_GEN=3 timeout 300 python scripts/spmv_probe.py --micro 0 --kernels 10 --reps 20 2>&1 | cut -c1-700 | tee gpurun_out/r02_spmv_probe_p_y3_$y3.txt; done
