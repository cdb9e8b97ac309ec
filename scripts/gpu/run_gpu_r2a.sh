#!/bin/bash
# round 2, call A: second-generation kernel bring-up (1 GPU): diagnostics, parity tests, A/B bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
echo "== v2_debug (product lib)"; timeout 240 python scripts/v2_debug.py > gpurun_out/r2a_debug.txt 2>&1; rc1=$?; echo "rc=$rc1" >> gpurun_out/r2a_debug.txt; tail -25 gpurun_out/r2a_debug.txt
echo "== v2_debug (slot dump build)"; MYRRIX_ALS_LIB=$PWD/scripts/_var/dbg.so timeout 240 python scripts/v2_debug.py > gpurun_out/r2a_debug_slot.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_debug_slot.txt; tail -3 gpurun_out/r2a_debug_slot.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log; tail -25 gpurun_out/r2a_pytest.log
if [ $rc1 -eq 0 ]; then
for rep in 1 2; do
  echo "v2:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2a_ab_v2.txt
  echo "v1:"; MYRRIX_ALS_V1=1 ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2a_ab_v1.txt
done
fi
