#!/bin/bash
mkdir -p gpurun_out
run() {
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-gmg 2>gpurun_out/u_err.txt | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'spmv_us', d['roofline']['avg_launch_us'], 'fx', d['history'][-1]['fx'])"
grep "tob\]" gpurun_out/u_err.txt | head -1
}
export TOB_VERBOSE=1
TOB_L2HINT=3 run "hint=3"
TOB_L2HINT=11 run "hint=11 persist=max"
TOB_L2HINT=11 TOB_L2_PERSIST_MB=48 run "hint=11 persist=48"
TOB_L2HINT=11 TOB_L2_PERSIST_MB=64 run "hint=11 persist=64"
TOB_L2HINT=15 TOB_L2_PERSIST_MB=48 run "hint=15 persist=48"
