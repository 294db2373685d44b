#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu 2>gpurun_out/o_cfg2_n${n}_err.txt | grep "^{" | tee gpurun_out/o_cfg2_n$n.json | cut -c1-200
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/o_cfg2_n${n}_err.txt | tail -3
done
