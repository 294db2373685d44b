#!/bin/bash
# final verification on one GPU: L2-hint A/B on the headline bench, full GPU suite, smoke, default bench, ncu launch list
mkdir -p gpurun_out
for m in 0 1 3 7; do
TOB_L2HINT=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-gmg 2>gpurun_out/t_hint${m}_err.txt | grep "^{" | tee gpurun_out/t_bench_hint$m.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('L2HINT=$m', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'spmv_us', d['roofline']['avg_launch_us'], d['config'].get('pcg_iterations_per_step'), 'fx', d['history'][-1]['fx'] if 'history' in d else None)"
done
TOB_L2HINT=7 timeout 300 python -m pytest tests/test_hex8_gpu.py -m gpu -x -q -k "solve or golden" 2>&1 | tail -2
echo "== full GPU suite"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
echo "== default bench"
timeout 900 python bench.py 2>gpurun_out/t_bench_err.txt | grep "^{" | tee gpurun_out/t_bench_default.json | cut -c1-600
tail -2 gpurun_out/t_bench_err.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-gmg > gpurun_out/t_ncu_bench.log 2>&1
tail -2 gpurun_out/t_ncu_bench.log | cut -c1-300
