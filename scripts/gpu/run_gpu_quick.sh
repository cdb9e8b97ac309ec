#!/bin/bash
# parity tests + quick device-resident bench lines (no e2e / cpu baseline)
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
./scripts/quick_bench.sh c3 3 2>&1 | tee gpurun_out/quick_c3.txt
./scripts/quick_bench.sh c2 5 2>&1 | tee gpurun_out/quick_c2.txt
echo "t=$(( $(date +%s)-S ))"
