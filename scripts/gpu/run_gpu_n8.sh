#!/bin/bash
# N-GPU check (gpurun --gpus N): sharded path vs single GPU on small shapes, then the C3 bench line at N.
N=${1:-8}
mkdir -p gpurun_out
S=$(date +%s)
./scripts/run_ranks.sh $N 200 gpurun_out/r02_n${N}_mgpu_check.log scripts/multi_gpu_check.py; tail -7 gpurun_out/r02_n${N}_mgpu_check.log
echo "t=$(( $(date +%s)-S ))"
./scripts/run_ranks.sh $N 400 gpurun_out/r02_n${N}_bench.log bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline; tail -1 gpurun_out/r02_n${N}_bench.log | cut -c1-2500
echo "t=$(( $(date +%s)-S ))"
