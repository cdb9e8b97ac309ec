#!/bin/bash
mkdir -p gpurun_out
echo "== v2_debug (phase profile build)"; MYRRIX_ALS_LIB=$PWD/scripts/_var/sprof.so timeout 240 python scripts/v2_debug.py 2>&1 | head -6
echo "== v2_debug"; timeout 240 python scripts/v2_debug.py > gpurun_out/r2h_debug.txt 2>&1; rc1=$?; echo "rc=$rc1" >> gpurun_out/r2h_debug.txt; tail -12 gpurun_out/r2h_debug.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -8 gpurun_out/r2h_pytest.log
if [ $rc1 -eq 0 ]; then bash scripts/gpu/run_gpu_ab.sh r2h nolock g2; fi
