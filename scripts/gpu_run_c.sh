#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== new tests (drop-in, reference MMA, multi-GPU)"
timeout 1500 python -m pytest tests/test_dropin_gpu.py tests/test_mma_ref.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/c_pytest.txt
echo "== bench N=2"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/c_bench2_err.txt | tee gpurun_out/c_bench2.json | cut -c1-1500
tail -5 gpurun_out/c_bench2_err.txt
