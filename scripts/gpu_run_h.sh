#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
echo "== 2-rank parity worker (P2P transport)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_worker.py > gpurun_out/h_mp.log 2>&1; echo "mp_worker rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/h_mp.log | tail -6
echo "== tobrun 2-rank drop-in test"
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -x -q -k "two_rank" 2>&1 | tail -3
for n in $N; do
echo "== bench N=$n (P2P)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu 2>gpurun_out/h_bench${n}_err.txt | grep "^{" | tee gpurun_out/h_bench$n.json | cut -c1-200
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/h_bench${n}_err.txt | tail -4
done
