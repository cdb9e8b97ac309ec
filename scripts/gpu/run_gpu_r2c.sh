#!/bin/bash
# round 2, call C: TMA-gather loader + in-place converters: correctness, then A/B against the register-gather build
mkdir -p gpurun_out
echo "== v2_debug"; timeout 240 python scripts/v2_debug.py > gpurun_out/r2c_debug.txt 2>&1; rc1=$?; echo "rc=$rc1" >> gpurun_out/r2c_debug.txt; tail -12 gpurun_out/r2c_debug.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -8 gpurun_out/r2c_pytest.log
if [ $rc1 -eq 0 ]; then
  echo "tma:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2c_ab_tma.txt
  for v in notma tma_x1 tma_pf4; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2c_ab_$v.txt; done
  echo "tma:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2c_ab_tma.txt
  echo "== wait profile"; timeout 300 python scripts/wait_profile.py c3p > gpurun_out/r2c_wait_profile.txt 2>&1; tail -24 gpurun_out/r2c_wait_profile.txt
fi
