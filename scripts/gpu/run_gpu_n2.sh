#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
./scripts/run_ranks.sh 2 200 gpurun_out/mgpu_check.log scripts/multi_gpu_check.py; tail -8 gpurun_out/mgpu_check.log
echo "t=$(( $(date +%s)-S ))"
./scripts/run_ranks.sh 2 300 gpurun_out/bench_n2.log bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e; tail -2 gpurun_out/bench_n2.log | cut -c1-1500
echo "t=$(( $(date +%s)-S ))"
