#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
for rep in 1 2; do
 echo auto;  ./scripts/quick_bench.sh c3 3 2>&1 | cut -c1-200
 echo mix8;  MYRRIX_ALS_MIX=8 ./scripts/quick_bench.sh c3 3 2>&1 | cut -c1-200
 echo mix4;  MYRRIX_ALS_MIX=4 ./scripts/quick_bench.sh c3 3 2>&1 | cut -c1-200
done
echo c2 auto; ./scripts/quick_bench.sh c2 5 2>&1 | cut -c1-200
echo c2 mix4; MYRRIX_ALS_MIX=4 ./scripts/quick_bench.sh c2 5 2>&1 | cut -c1-200
echo "t=$(( $(date +%s)-S ))"
