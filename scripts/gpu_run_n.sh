#!/bin/bash
mkdir -p gpurun_out
TOB_BENCH_NOPROFILE=1 timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-gmg 2>/dev/null | grep "^{" | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('noprofile', b['value'], b['ms_per_step'], b['e2e']['value'])"
TOB_PCG_UNFUSED=1 TOB_BENCH_NOPROFILE=1 timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-gmg 2>/dev/null | grep "^{" | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('noprofile unfused', b['value'], b['ms_per_step'])"
