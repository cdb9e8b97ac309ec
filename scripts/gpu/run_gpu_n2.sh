#!/bin/bash
# 2-GPU check (gpurun --gpus 2): sharded path vs single GPU (synthetic shards, then host blocks + distributed
# transpose), with the P2P row push and with the NCCL all-gather fallback; then the C3 bench line at N = 2.
mkdir -p gpurun_out
S=$(date +%s)
./scripts/run_ranks.sh 2 200 gpurun_out/r02_n2_mgpu_check.log scripts/multi_gpu_check.py; tail -8 gpurun_out/r02_n2_mgpu_check.log
echo "t=$(( $(date +%s)-S ))"
MYRRIX_ALS_NO_P2P=1 ./scripts/run_ranks.sh 2 200 gpurun_out/r02_n2_mgpu_check_nop2p.log scripts/multi_gpu_check.py; tail -6 gpurun_out/r02_n2_mgpu_check_nop2p.log
echo "t=$(( $(date +%s)-S ))"
./scripts/run_ranks.sh 2 400 gpurun_out/r02_n2_bench_n2.log bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline; tail -2 gpurun_out/r02_n2_bench_n2.log | cut -c1-3000
echo "t=$(( $(date +%s)-S ))"
MYRRIX_ALS_NO_P2P=1 ./scripts/run_ranks.sh 2 300 gpurun_out/r02_n2_bench_n2_nop2p.log bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity; tail -2 gpurun_out/r02_n2_bench_n2_nop2p.log | cut -c1-1500
echo "t=$(( $(date +%s)-S ))"
