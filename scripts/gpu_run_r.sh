#!/bin/bash
# 2-GPU: replicated coarse multigrid levels + multi-rank result container; N=2 bench with both solver legs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_dropin_gpu.py::test_two_rank_run_matches_single tests/test_hex8_gpu.py -m gpu -x -q -k "multi or two_rank or gmg or GMG or mg" 2>&1 | tail -8
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r_cfg2_n${n}_err.txt | grep "^{" | tee gpurun_out/r_cfg2_n$n.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d.get('gmg_solver'))"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r_cfg2_n${n}_err.txt | tail -5
