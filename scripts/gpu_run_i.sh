#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 3 --warmup 3 --workload cfg3 --no-cpu 2>gpurun_out/i_cfg3_n1_err.txt | grep "^{" | tee gpurun_out/i_cfg3_n1.json | cut -c1-300
  tail -3 gpurun_out/i_cfg3_n1_err.txt
else
  for n in $N; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 --workload cfg3 --no-cpu 2>gpurun_out/i_cfg3_n${n}_err.txt | grep "^{" | tee gpurun_out/i_cfg3_n$n.json | cut -c1-300
    grep -v "^\*\|OMP_NUM\|^$" gpurun_out/i_cfg3_n${n}_err.txt | tail -4
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu 2>gpurun_out/i_cfg2_n${n}_err.txt | grep "^{" | tee gpurun_out/i_cfg2_n$n.json | cut -c1-300
    grep -v "^\*\|OMP_NUM\|^$" gpurun_out/i_cfg2_n${n}_err.txt | tail -4
  done
fi
