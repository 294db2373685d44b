#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_worker.py > gpurun_out/f_mp.log 2>&1; echo "mp_worker rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/f_mp.log | tail -15
