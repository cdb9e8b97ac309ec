#!/bin/bash
# c3 regression check of the plain kernel variant; c5p with stash mode off / on at 5 steps
mkdir -p gpurun_out
BENCH_TIMEOUT=100 bash scripts/quick_bench.sh c3 3 2>&1 | tee gpurun_out/s3e_c3.txt
MYRRIX_ALS_STASH=0 BENCH_TIMEOUT=60 bash scripts/quick_bench.sh c5p 5 2>&1 | tee gpurun_out/s3e_c5p_nostash.txt
BENCH_TIMEOUT=60 bash scripts/quick_bench.sh c5p 5 2>&1 | tee gpurun_out/s3e_c5p_stash.txt
timeout 60 python -m pytest tests/test_parity_gpu.py -x -q -k "stash or split" 2>&1 | tail -2
