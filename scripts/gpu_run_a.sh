#!/bin/bash
# First GPU pass: parity tests, FP64 micro-benchmarks, SpMV GB/s, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/a_pytest.txt
echo "== probe"
timeout 600 python scripts/spmv_probe.py 2>&1 | tee gpurun_out/a_probe.txt
echo "== ncu launch list (256^3, zmarch ty8 + ty16)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches.csv \
   python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels 2,3,4 --micro 0 > gpurun_out/a_ncu1.log 2>&1
echo "== ncu full capture of the z-march kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zmarch -s 4 -c 2 -o gpurun_out/a_zmarch12 \
   python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels 4 --micro 0 > gpurun_out/a_ncu2.log 2>&1
ls -la gpurun_out
