#!/bin/bash
# round 2, call D: producers gather asynchronously for themselves (cp.async into the ring, in-place conversion)
mkdir -p gpurun_out
echo "== v2_debug"; timeout 240 python scripts/v2_debug.py > gpurun_out/r2d_debug.txt 2>&1; rc1=$?; echo "rc=$rc1" >> gpurun_out/r2d_debug.txt; tail -4 gpurun_out/r2d_debug.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; tail -8 gpurun_out/r2d_pytest.log
if [ $rc1 -eq 0 ]; then
  echo "async:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2d_ab_async.txt
  for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2d_ab_$v.txt; done
  echo "async:"; ./scripts/quick_bench.sh c3 3 2>&1 | tee -a gpurun_out/r2d_ab_async.txt
  echo "== wait profile"; timeout 300 python scripts/wait_profile.py c3p > gpurun_out/r2d_wait_profile.txt 2>&1; tail -26 gpurun_out/r2d_wait_profile.txt
fi
