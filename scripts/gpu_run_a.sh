#!/bin/bash
# GPU pass: parity tests, FP64 micro-benchmarks, SpMV GB/s, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/a_pytest.txt
fi
echo "== probe"
timeout 600 python scripts/spmv_probe.py --micro ${MICRO:-1} 2>&1 | tee gpurun_out/a_probe.txt
echo "== ncu launch list (256^3)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches.csv \
   python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels 2,3,4,5 --micro 0 > gpurun_out/a_ncu1.log 2>&1
for k in ${NCU_KERNELS:-4 5}; do
echo "== ncu full capture of z-march kernel variant $k"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hex8_zmarch -s 4 -c 1 -o gpurun_out/a_zmarch_k$k \
   python scripts/spmv_probe.py --sizes 256x256x256 --reps 3 --kernels $k --micro 0 > gpurun_out/a_ncu2_k$k.log 2>&1
done
ls -la gpurun_out
