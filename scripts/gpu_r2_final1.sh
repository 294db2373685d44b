#!/bin/bash
# final single-GPU pass of the round: parity suite (the 12-minute CPU-bound cfg2 reference-solver test is deselected here; the
# driver runs it), smoke, default bench line, 20-step main arm, TF32 smoother comparison, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_hex8_gpu.py::test_iteration1_compliance_cfg2_vs_cpu_reference_solver 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2>gpurun_out/r02_bench_final_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_final_n1.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')},'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['avg_launch_us'],'its',d['config']['pcg_iterations_per_step'])
print('spmv256',d['spmv_256']['frac'],'gmg',d['gmg_solver']['value'],d['gmg_solver']['pcg_iterations_per_step'],'cfg3',d['cfg3_gmg'].get('value'),d['cfg3_gmg'].get('pcg_iterations_per_step'))
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline'].get('seconds_per_step'),'clocks',d['clocks'])
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-gmg --no-cfg3 --no-spmv256 2>/dev/null | grep "^{" > gpurun_out/r02_bench_final_n1_20steps.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_final_n1_20steps.json').read()); print('20 steps:', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])"
for tf in 0 1; do TOB_GMG_TF32=$tf TOB_GMG_TIMING=1 timeout 600 python scripts/gmg_growth_probe.py cfg3 8 2:3 2>&1 | grep -e "V-cycle" -e workload | tail -2 | cut -c1-600; done | tee gpurun_out/r02_tf32_smoother_final.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_bench_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-gmg --no-cfg3 --no-spmv256 > gpurun_out/r02_ncu_launches.log 2>&1
