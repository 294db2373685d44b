#!/bin/bash
# N GPUs (default 4): cfg2 bench (Jacobi + GMG legs) and cfg3 with the GMG solver
mkdir -p gpurun_out
n=${NG:-4}
for w in "cfg2 jacobi" "cfg3 gmg"; do
set -- $w
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu --workload $1 --solver $2 2>gpurun_out/s_$1_n${n}_err.txt | grep "^{" | tee gpurun_out/s_$1_n$n.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['config'].get('pcg_iterations_per_step'), d.get('gmg_solver'))"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s_$1_n${n}_err.txt | tail -3
done
