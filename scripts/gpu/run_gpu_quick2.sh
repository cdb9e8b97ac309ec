#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
echo base; ./scripts/quick_bench.sh c3 3 2>&1 | tee gpurun_out/quick_c3.txt
for v in "$@"; do echo "$v:"; MYRRIX_ALS_LIB=$PWD/scripts/_var/$v.so ./scripts/quick_bench.sh c3 3 2>&1 | tee gpurun_out/quick_c3_$v.txt; done
./scripts/quick_bench.sh c2 5 2>&1 | tee gpurun_out/quick_c2.txt
echo "t=$(( $(date +%s)-S ))"
