#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/g_bench1_err.txt | tee gpurun_out/g_bench1.json | cut -c1-400
  tail -3 gpurun_out/g_bench1_err.txt
else
  for n in 4 8; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 3 --warmup 3 2>gpurun_out/g_bench${n}_err.txt | tee gpurun_out/g_bench$n.json | cut -c1-300
    grep -v "^\*\|OMP_NUM\|^$" gpurun_out/g_bench${n}_err.txt | tail -5
  done
fi
